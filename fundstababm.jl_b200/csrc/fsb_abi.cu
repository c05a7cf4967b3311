// fsb_abi.cu -- host side of libfsb_b200.so: the C ABI of include/fsb.h.
// Owns device memory, the stream, the launch schedule of the per-step kernel and the NCCL
// communicator (libnccl is dlopen'ed on first use so the library has no link-time dependency).
// There is no CPU fallback: every compute entry point needs a CUDA device.
#include <cuda.h>  // CUtensorMap and its enums only: the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fsb.h"
#include "fsb_aux.cuh"
#include "fsb_state.h"
#include "fsb_step.cuh"
#include "fsb_passx.cuh"

namespace fsb {  // the FSB_BIG build of the step kernels (fsb_step_big.cu)
__global__ void fsb_big_events1_kernel(const DevState s, const StepArgs a, const int launch_parity);
__global__ void fsb_big_pass_kernel(const DevState s, const StepArgs a, const int launch_parity, const __grid_constant__ CUtensorMap tm_h,
                                    const __grid_constant__ CUtensorMap tm_h16);
__global__ void fsb_big_events2_kernel(const DevState s, const StepArgs a, const int launch_parity);
}  // namespace fsb

namespace fsb {  // the FSB_FIXED build of the event kernels (fsb_step_fixed.cu): exactly the default Params shape
__global__ void fsb_fix_events1_kernel(const DevState s, const StepArgs a, const int launch_parity);
__global__ void fsb_fix_choice_kernel(const DevState s, const StepArgs a, const int launch_parity);
__global__ void fsb_fix_events2_kernel(const DevState s, const StepArgs a, const int launch_parity);
constexpr int kFixedK = 250, kFixedM = 250, kFixedN = 1000, kFixedCap = 80, kFixedSlots = 8, kFixedW = 100;  // == fsb_step.cuh under FSB_FIXED
}  // namespace fsb

namespace fsb {  // fsb_facts.cu
int facts_max_lag();
int facts_max_sel();
long long facts_doubles_per_rep(int M, int maxlag);
cudaError_t facts_launch(const DevState &d, cudaStream_t st, long long rep_first, int n_reps, int maxlag, double pct, double *base);
}  // namespace fsb

using namespace fsb;

constexpr int kMaxBatches = 8;  // replication batches (streams) an engine can run concurrently

// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int32_t fail(int32_t code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) return fail(FSB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// ---- NCCL through dlopen -------------------------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int32_t nccl_load()
{
    if (g_nccl.lib) return FSB_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(FSB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                             \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));         \
    if (!g_nccl.field) return fail(FSB_ERR_NCCL, "libnccl lacks symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommInitAll, "ncclCommInitAll");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return FSB_OK;
}
#define NC(expr)                                                                                        \
    do {                                                                                                \
        ncclResult_t r_ = (expr);                                                                       \
        if (r_ != ncclSuccess) return fail(FSB_ERR_NCCL, "%s: %s", #expr, g_nccl.GetErrorString(r_));   \
    } while (0)

// ------------------------------------------------------------------------------------------------
struct fsb_engine {
    CUtensorMap tm_h{};  // 2-D view of the holdings of ALL replications: [R * K4 rows][Mp], box = 8 rows x one stage (pass kernel)
    CUtensorMap tm_h16{};  // the same view with a box of 16 rows x half a stage row (full_pass_tma16)
    DevState d{};
    fsb_config cfg{};
    std::vector<DevPoint> points;
    std::vector<void *> allocs;
    // FSB_REDZONE=1 (debug): every device buffer sits between two guard bands filled with a pattern; fsb_debug_redzones
    // counts the buffers whose bands were overwritten (compute-sanitizer is not available on every pool)
    bool redzone = false;
    std::vector<std::pair<unsigned char *, size_t>> zones;  // (user pointer, user bytes)
    int64_t device_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // replication batches: each runs its kernels of a step on its own stream, so that the memory-bound
    // holdings pass of one batch overlaps the latency-bound event kernels of another
    int block_e2 = 0;
    int sms = 1;
    int n_batches = 1;
    std::vector<cudaStream_t> bstream;
    std::vector<cudaEvent_t> bjoin;
    cudaEvent_t fork = nullptr;
    int occ[3] = {1, 1, 1};
    int grid[3] = {0, 0, 0}, block = 256, block_pass = 256, smem_bytes[3] = {0, 0, 0}, init_smem = 0;
    int launch_parity = 0;
    bool fast = false;
    bool usex = false;  // the pass reads the packed mirror (fsb_passx_kernel)
    bool fixed = false; // the event / choice kernels come from the fixed-shape build (fsb_step_fixed.cu)
    int grid_x = 0;
    bool log_built = false, state_set = false;
    int64_t t_done = 0;  // last simulated t (0 = none since init/upload)
    int64_t words_per_point = 0;
    unsigned long long *stats_dev = nullptr;
    ncclComm_t comm = nullptr;
    int n_ranks = 1;
    // last-run timing
    double last_total_ms = 0.0, last_pass_ms = 0.0, last_facts_ms = 0.0;
    int64_t last_launches = 0, last_pass_launches = 0;
    std::vector<cudaEvent_t> kev;  // event pairs around the pass-kernel launches of the last run
    size_t kev_used = 0;
    // host mirrors of the per-replication pointer tables
    std::vector<const double *> z_mkt, z_stock, u_review;
    std::vector<const fsb_tape_rec *> ev_tape;
    std::vector<long long> ev_n;
    std::vector<int> trace_slot;
    int n_trace_slots = 0;
    // device pointer tables
    const double **z_mkt_d = nullptr, **z_stock_d = nullptr, **u_review_d = nullptr;
    const fsb_tape_rec **ev_tape_d = nullptr;
    long long *ev_n_d = nullptr;
    int *trace_slot_d = nullptr;
    double *st_zmkt = nullptr, *st_zstock = nullptr, *st_urev = nullptr;
};

template <class T>
static int32_t dalloc(fsb_engine *e, T **out, size_t count, bool zero = true)
{
    void *p = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    if (e->redzone) {
        constexpr size_t kBand = 256;
        const size_t padded = (bytes + 255) & ~static_cast<size_t>(255);
        CU(cudaMalloc(&p, padded + 2 * kBand));
        CU(cudaMemset(p, 0xA5, padded + 2 * kBand));
        unsigned char *u = static_cast<unsigned char *>(p) + kBand;
        if (zero) CU(cudaMemset(u, 0, bytes));
        e->allocs.push_back(p);
        e->zones.emplace_back(u, bytes);
        e->device_bytes += static_cast<int64_t>(padded + 2 * kBand);
        *out = reinterpret_cast<T *>(u);
        return FSB_OK;
    }
    CU(cudaMalloc(&p, bytes));
    if (zero) CU(cudaMemset(p, 0, bytes));
    e->allocs.push_back(p);
    e->device_bytes += static_cast<int64_t>(bytes);
    *out = static_cast<T *>(p);
    return FSB_OK;
}
#define TRY(expr)                    \
    do {                             \
        int32_t rc_ = (expr);        \
        if (rc_ != FSB_OK) return rc_; \
    } while (0)

extern "C" int32_t fsb_version(void) { return FSB_VERSION; }
extern "C" const char *fsb_last_error(void) { return g_err.c_str(); }

extern "C" int64_t fsb_stats_words_per_point(const fsb_engine *e)
{
    return e ? e->words_per_point : 0;
}
extern "C" int64_t fsb_device_bytes(const fsb_engine *e) { return e ? e->device_bytes : 0; }
extern "C" int64_t fsb_next_day(const fsb_engine *e) { return e ? (e->t_done > 0 ? e->t_done + 1 : e->d.W + 2) : 0; }

extern "C" uint64_t fsb_tape_key(int32_t site, int64_t t, int64_t a, int64_t b)
{
    return (static_cast<uint64_t>(site & 0xff) << 56) | (static_cast<uint64_t>(t & 0xffffff) << 32) |
           (static_cast<uint64_t>(a & 0xffff) << 16) | static_cast<uint64_t>(b & 0xffff);
}

static int32_t check_params(const fsb_params *p, int32_t n_points)
{
    if (!p || n_points < 1) return fail(FSB_ERR_BAD_PARAMS, "need at least one Params point");
    for (int i = 0; i < n_points; ++i) {
        const fsb_params &q = p[i];
        if (q.bigk < 1 || q.bigm < 1 || q.bign < q.bigk) return fail(FSB_ERR_BAD_PARAMS, "point %d: need bign >= bigk >= 1, bigm >= 1", i);
        if (!(0 < q.perf_lo && q.perf_lo <= q.perf_hi && q.perf_hi < q.bigt)) return fail(FSB_ERR_BAD_PARAMS, "point %d: need 0 < perfwindow[1] <= perfwindow[end] < bigt", i);
        if (q.psize_lo < 1 || q.psize_lo > q.psize_hi) return fail(FSB_ERR_BAD_PARAMS, "point %d: bad portfsizerange", i);
        if (q.cap_lo < 0 || q.cap_lo > q.cap_hi) return fail(FSB_ERR_BAD_PARAMS, "point %d: bad invcaprange", i);
        if (q.n_impact < 1 || q.n_vol < 1 || !q.impact_values || !q.vol_values) return fail(FSB_ERR_BAD_PARAMS, "point %d: empty impactrange/stockvolrange", i);
        if (!(q.reviewprobability >= 0.0 && q.reviewprobability <= 1.0)) return fail(FSB_ERR_BAD_PARAMS, "point %d: reviewprobability is not a probability", i);
        if (q.bigm != p[0].bigm || q.bign != p[0].bign || q.bigt != p[0].bigt || q.bigk != p[0].bigk || q.perf_hi != p[0].perf_hi)
            return fail(FSB_ERR_BAD_PARAMS, "point %d: shapes (bigm,bign,bigt,bigk,perfwindow[end]) must agree across sweep points", i);
        if (q.bign > 65535 || q.bigm > 65535 || q.bigt >= (1 << 24)) return fail(FSB_ERR_BAD_PARAMS, "point %d: shape exceeds tape key range", i);
    }
    return FSB_OK;
}

extern "C" int32_t fsb_create(const fsb_params *pts, int32_t n_points, const fsb_config *cfg, fsb_engine **out)
{
    if (!out || !cfg) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    *out = nullptr;
    TRY(check_params(pts, n_points));
    if (cfg->n_reps < 1) return fail(FSB_ERR_BAD_PARAMS, "n_reps must be >= 1");
    if (cfg->reps_per_point > 0 && (cfg->rep_offset + cfg->n_reps + cfg->reps_per_point - 1) / cfg->reps_per_point > n_points)
        return fail(FSB_ERR_BAD_PARAMS, "replication range maps beyond the %d sweep points", n_points);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(FSB_ERR_CUDA, "no CUDA device: libfsb_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FSB_ERR_BAD_PARAMS, "device %d out of range (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    fsb_engine *e = new fsb_engine();
    e->cfg = *cfg;
    e->redzone = getenv("FSB_REDZONE") && atoi(getenv("FSB_REDZONE")) > 0;
    DevState &d = e->d;
    const fsb_params &p0 = pts[0];
    d.K = static_cast<int>(p0.bigk); d.M = static_cast<int>(p0.bigm); d.N = static_cast<int>(p0.bign);
    d.T = static_cast<int>(p0.bigt); d.W = static_cast<int>(p0.perf_hi);
    d.Mp = (d.M + 127) & ~127;  // 1 KiB row granules: no tails or masks in the 8-lane loops, rows start on 128-byte lines
    d.keep_history = cfg->keep_history ? 1 : 0;
    d.hist_cols = d.keep_history ? d.T : 2 * (d.W + 1);
    d.mkt_len = d.keep_history ? d.T : 2;
    d.R = cfg->n_reps; d.rep_offset = cfg->rep_offset; d.reps_per_point = cfg->reps_per_point; d.n_points = n_points;
    d.h_stride = ((static_cast<long long>((d.K + 3) & ~3) * d.Mp + 15) / 16) * 16;
    d.p_stride = static_cast<long long>(d.hist_cols) * d.Mp;
    // event capacity: reviewers ~ Binomial(N, p); mean + 12 sd + 16
    double pmax = 0.0;
    for (int i = 0; i < n_points; ++i) pmax = std::max(pmax, pts[i].reviewprobability);
    const double mean = d.N * pmax;
    int cap = cfg->event_capacity > 0 ? cfg->event_capacity : static_cast<int>(std::ceil(mean + 12.0 * std::sqrt(mean * (1.0 - pmax)) + 16.0));
    cap = std::min(std::max(cap, 8), d.N);
    cap = std::max(cap, std::min(d.N, 8));
    d.cap_ev = (cap + 7) & ~7;
    d.cap_log = cfg->log_capacity > 0 ? std::max(cfg->log_capacity, d.K + 1) : d.K + 1024;
    d.cap_trace = cfg->trace_capacity > 0 ? cfg->trace_capacity : 1 << 20;
    e->block = cfg->block_threads > 0 ? ((cfg->block_threads + 31) / 32) * 32 : kStepThreads;
    if (e->block > kStepThreads) e->block = kStepThreads;
    if (d.K >= 32768) { delete e; return fail(FSB_ERR_BAD_PARAMS, "bigk must be < 32768 (16-bit fund ids in shared memory)"); }
    // shared memory of the three kernels of a step.  events-1 / events-2: the two price columns + a few
    // order / probability rows (more rows spill to a global scratch); pass: the largest horizon-column
    // chunk (2, 6, 10 or 14 columns) that leaves room for the wanted number of CTAs per SM
    int smem_max = 0, smem_sm = 0, sms = 0;
    CU(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
    CU(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, cfg->device));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
    e->sms = sms;
    int nw = e->block / 32;
    // events-2 may run wider CTAs than events-1 (one warp per horizon column in its rank phase)
    e->block_e2 = (cfg->block_threads > 0) ? e->block : kE2Threads;
    int nw2 = e->block_e2 / 32;
    e->block_pass = (d.K <= 256 && d.Mp == 256 && !getenv("FSB_NO_TMA")) ? kPassThreads : e->block;
    if (const char *ev = getenv("FSB_PASS_THREADS")) e->block_pass = std::max(32, std::min(kPassThreads, (atoi(ev) / 32) * 32));
    int nwp = e->block_pass / 32;
    d.ring_stages = (d.K <= 256 && d.Mp == 256) ? kRingStages : 0;
    if (getenv("FSB_NO_TMA")) d.ring_stages = 0;
    // big shapes: even the leanest events-1 layout (3 column slots) does not fit in shared memory
    d.big = (make_layout(d.K, d.Mp, d.N, d.cap_ev, 3, nw).bytes + 64 > smem_max || getenv("FSB_FORCE_BIG")) ? 1 : 0;
    if (d.big && cfg->block_threads <= 0) { e->block = e->block_e2 = kBigStepThreads; nw = nw2 = kBigStepThreads / 32; }
    // big shapes: ring_stages > 0 selects the tiled pass (big_pass_tiled; FSB_BIG_TILED=0: whole columns from the CTA's global buffer)
    if (d.big) {
        d.ring_stages = (getenv("FSB_BIG_TILED") && atoi(getenv("FSB_BIG_TILED")) == 0) ? 0 : 2;
        const int dflt = d.ring_stages > 0 ? kBigPassThreads : kPassThreads;  // (tiled: the last warp only feeds the column ring)
        e->block_pass = getenv("FSB_BIG_PASS_THREADS") ? std::max(64, std::min(dflt, (atoi(getenv("FSB_BIG_PASS_THREADS")) / 32) * 32)) : dflt;
    }
    d.pass_parts = 1;
    d.pass_chunks = 1;
    if (d.big && d.ring_stages > 0) {  // up to 9 work units per replication (whole row blocks each): an even finish of the pass kernel's persistent CTAs
        const int rpb = 2 * kBigR * (e->block_pass / 32 - kBigProd), nblocks = (((d.K + 3) & ~3) + rpb - 1) / rpb;  // (config 4: 36 blocks of 56 rows)
        d.pass_parts = std::max(1, std::min(getenv("FSB_PASS_PARTS") ? atoi(getenv("FSB_PASS_PARTS")) : 9, nblocks));
        d.pass_chunks = std::max(1, std::min(getenv("FSB_PASS_CHUNKS") ? atoi(getenv("FSB_PASS_CHUNKS")) : 3, 8));
    }
    int want_ctas = cfg->blocks_per_sm > 0 ? cfg->blocks_per_sm : (d.ring_stages ? 2 : 4);
    if (const char *ev = getenv("FSB_CTAS_PER_SM")) want_ctas = std::max(1, atoi(ev));
    const int budget = std::min(smem_max, smem_sm / want_ctas - 1024);
    d.cmax = kMaxCols;
    if (const char *ev = getenv("FSB_CMAX")) d.cmax = std::min(kMaxCols, std::max(6, ((atoi(ev) + 2) / 4) * 4 - 2));
    // (a chunk after the first holds cmax - 2 columns: cmax stays >= 6)
    nwp = e->block_pass / 32;
    while (!d.big && make_layout(d.K, d.Mp, d.N, d.cap_ev, d.cmax + 2, nwp, d.ring_stages).bytes + 64 > budget && d.cmax > 6) d.cmax -= 4;
    d.slots_a = 2 + 6;
    d.slots_c = 2 + std::max(nw2, 6);
    if (const char *ev = getenv("FSB_EVENT_ROWS")) { d.slots_a = 2 + std::max(1, atoi(ev)); d.slots_c = 2 + std::max(nw2, atoi(ev)); }
    while (!d.big && make_layout(d.K, d.Mp, d.N, d.cap_ev, d.slots_a, nw).bytes + 64 > smem_max && d.slots_a > 3) --d.slots_a;
    while (!d.big && make_layout(d.K, d.Mp, d.N, d.cap_ev, d.slots_c, nw2).bytes + 64 > smem_max && d.slots_c > 2 + nw2) --d.slots_c;
    d.kc = std::min(d.W, std::max(std::max(32, 2 * kMaxCols), d.cap_ev / 2));  // (distinct horizons of a step's cash holders)
    d.hi_stride = hand_ints(d.K, d.cap_ev);
    e->init_smem = (d.W + 2 + d.K) * 8;
    if (e->init_smem > 48 * 1024) CU(cudaFuncSetAttribute(fsb_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, e->init_smem));
    {
        const int slots[3] = {d.slots_a, d.cmax + 2, d.slots_c};
        const void *kern_std[3] = {(const void *)fsb_events1_kernel, (const void *)fsb_pass_kernel, (const void *)fsb_events2_kernel};
        const void *kern_big[3] = {(const void *)fsb_big_events1_kernel, (const void *)fsb_big_pass_kernel, (const void *)fsb_big_events2_kernel};
        const void *const *kern = d.big ? kern_big : kern_std;
        d.big_stride = 0;
        for (int k = 0; k < 3; ++k) {
            const int blk = (k == 1) ? e->block_pass : (k == 2) ? e->block_e2 : e->block;
            const SmemLayout L = make_layout(d.K, d.Mp, d.N, d.cap_ev, slots[k], blk / 32, k == 1 ? d.ring_stages : 0, d.big != 0);
            d.big_stride = std::max(d.big_stride, L.big_bytes);
            if (L.bytes + 64 > smem_max) { delete e; return fail(FSB_ERR_BAD_PARAMS, "shapes need %d bytes of shared memory per CTA (> %d)", L.bytes, smem_max); }
            e->smem_bytes[k] = L.bytes;
            CU(cudaFuncSetAttribute(kern[k], cudaFuncAttributeMaxDynamicSharedMemorySize, L.bytes));
#if FSB_SPLIT_CHOICE
            if (k == 2 && !d.big) CU(cudaFuncSetAttribute((const void *)fsb_choice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.bytes));
#endif
            int occ = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern[k], blk, L.bytes));
            if (occ < 1) { delete e; return fail(FSB_ERR_CUDA, "step kernel %d cannot be resident (smem %d)", k, L.bytes); }
            if (cfg->blocks_per_sm > 0 && k == 1) occ = std::min(occ, cfg->blocks_per_sm);
            if (const char *ev = getenv(k == 1 ? "FSB_PASS_CTAS" : "FSB_EVENT_CTAS")) occ = std::min(occ, std::max(1, atoi(ev)));
            e->occ[k] = occ;
            e->grid[k] = static_cast<int>(std::min<long long>(d.R * (k == 1 ? d.pass_parts * d.pass_chunks : 1), static_cast<long long>(sms) * occ));
        }
    }
    d.scan_choice = (getenv("FSB_SCAN_CHOICE") && atoi(getenv("FSB_SCAN_CHOICE")) > 0) ? 1 : 0;
    d.smem_sort = (getenv("FSB_SMEM_SORT") && atoi(getenv("FSB_SMEM_SORT")) > 0) ? 1 : 0;
    d.pass16 = (getenv("FSB_PASS16") && atoi(getenv("FSB_PASS16")) == 0) ? 0 : 1;
    d.pass_ahead = (getenv("FSB_PASS_AHEAD") && atoi(getenv("FSB_PASS_AHEAD")) == 0) ? 0 : 1;
    e->fast = (d.K <= 256);
    // the default Params shape exactly: the event and choice kernels of the fixed-shape build (FSB_NO_FIXED=1: generic build)
    e->fixed = (!d.big && d.K == kFixedK && d.M == kFixedM && d.N == kFixedN && d.W == kFixedW && d.cap_ev == kFixedCap && d.slots_a == kFixedSlots &&
                d.slots_c == kFixedSlots && e->block == kStepThreads && e->block_e2 == kE2Threads && kE2Threads == kStepThreads && !getenv("FSB_NO_FIXED"));
    if (e->fixed) {
        const void *kf[3] = {(const void *)fsb_fix_events1_kernel, (const void *)fsb_fix_choice_kernel, (const void *)fsb_fix_events2_kernel};
        const int sm[3] = {e->smem_bytes[0], e->smem_bytes[2], e->smem_bytes[2]};
        for (int k = 0; k < 3; ++k) {
            int occ = 0;
            if (cudaFuncSetAttribute(kf[k], cudaFuncAttributeMaxDynamicSharedMemorySize, sm[k]) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf[k], e->block, sm[k]) != cudaSuccess || occ < e->occ[k == 0 ? 0 : 2]) {
                (void)cudaGetLastError();
                e->fixed = false;  // (cannot be as resident as the generic build: keep that one)
            }
        }
    }
    // batches: with more than one, every kernel takes only a share of each SM so that kernels of different
    // batches are co-resident (pass: 2 CTAs per SM, events: 3)
    e->n_batches = 1;  // (FSB_BATCHES > 1: kernels of different batches overlap; measured gain 3 % at most)
    if (const char *ev = getenv("FSB_BATCHES")) e->n_batches = std::max(1, std::min(kMaxBatches, atoi(ev)));
    if (e->n_batches > d.R || d.big) e->n_batches = 1;
    if (e->n_batches > 1 && (getenv("FSB_SHARE_EVENTS") || getenv("FSB_SHARE_PASS"))) {
        int share[3] = {3, 2, 3};
        if (const char *ev = getenv("FSB_SHARE_EVENTS")) share[0] = share[2] = std::max(1, atoi(ev));
        if (const char *ev = getenv("FSB_SHARE_PASS")) share[1] = std::max(1, atoi(ev));
        for (int k = 0; k < 3; ++k) e->grid[k] = sms * std::min(e->occ[k], share[k]);
    }

    // points + tables
    std::vector<double> tables;
    for (int i = 0; i < n_points; ++i) {
        const fsb_params &q = pts[i];
        DevPoint dp{};
        dp.marketstartval = q.marketstartval; dp.drift = q.drift; dp.marketvol = q.marketvol;
        dp.betamean = q.betamean; dp.betastd = q.betastd; dp.stockstartval = q.stockstartval;
        dp.reviewprob = q.reviewprobability; dp.thresholdmean = q.thresholdmean; dp.thresholdstd = q.thresholdstd;
        dp.perf_lo = static_cast<int>(q.perf_lo); dp.perf_hi = static_cast<int>(q.perf_hi);
        dp.cap_lo = static_cast<int>(q.cap_lo); dp.cap_hi = static_cast<int>(q.cap_hi);
        dp.psize_lo = static_cast<int>(q.psize_lo); dp.psize_hi = static_cast<int>(std::min<int64_t>(q.psize_hi, 1 << 30));
        dp.n_impact = static_cast<int>(q.n_impact); dp.n_vol = static_cast<int>(q.n_vol);
        dp.impact_off = static_cast<int>(tables.size());
        tables.insert(tables.end(), q.impact_values, q.impact_values + q.n_impact);
        dp.vol_off = static_cast<int>(tables.size());
        tables.insert(tables.end(), q.vol_values, q.vol_values + q.n_vol);
        e->points.push_back(dp);
    }
    auto cleanup = [&](int32_t rc) { fsb_destroy(e); return rc; };
#define A(field, count) do { int32_t rc_ = dalloc(e, &d.field, (count)); if (rc_ != FSB_OK) return cleanup(rc_); } while (0)
    const size_t R = static_cast<size_t>(d.R);
    A(H, R * d.h_stride);
    A(P, R * d.p_stride);
    if (d.ring_stages > 0) {  // the pass kernel fetches a stage (8 rows x kStageRowBytes) with ONE tensor copy
        if (d.h_stride != static_cast<long long>((d.K + 3) & ~3) * d.Mp) return cleanup(fail(FSB_ERR_BAD_PARAMS, "holdings stride is not K4 * Mp"));
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return cleanup(fail(FSB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver"));
        const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(d.Mp), static_cast<cuuint64_t>(R) * ((d.K + 3) & ~3)};
        const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(d.Mp) * 8};
        // (default shape: 8 rows x kStageRowBytes; big shapes: the trip of big_pass_tiled)
        const cuuint32_t box[2] = {static_cast<cuuint32_t>((d.big ? kBigStageRowBytes : kStageRowBytes) / 8), static_cast<cuuint32_t>(d.big ? kBigStageRows : 8)}, estr[2] = {1, 1};
        const CUresult cr = reinterpret_cast<encode_fn>(fn)(&e->tm_h, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d.H, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) return cleanup(fail(FSB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(cr)));
        e->tm_h16 = e->tm_h;
        if (!d.big) {
            const cuuint32_t box16[2] = {static_cast<cuuint32_t>(kStageBytes / 16 / 8), 16};
            const CUresult cr2 = reinterpret_cast<encode_fn>(fn)(&e->tm_h16, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d.H, gdim, gstr, box16, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr2 != CUDA_SUCCESS) return cleanup(fail(FSB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(cr2)));
        }
    }
    // the packed mirror of the holdings (fsb_passx.cuh), default shape only.  OPT-IN (FSB_PASS_PACKED=1): measured on a B200 it is
    // bit-identical to the dense TMA-fed pass and moves 0.4x its bytes, but runs at 0.77x its speed -- both are bound by
    // shared-memory wavefronts, not by HBM -- so the dense pass stays the default.
    e->usex = (d.K <= 256 && d.Mp == 256 && !d.big && d.kc <= kXKc && d.ring_stages > 0 && getenv("FSB_PASS_PACKED") && atoi(getenv("FSB_PASS_PACKED")) > 0);
    d.xcols = kXCols;
    if (const char *ev = getenv("FSB_XCOLS")) d.xcols = std::min(kXCols, std::max(4, (atoi(ev) / 4) * 4));
    if (e->usex) {
        A(X, R * ((d.K + 3) & ~3) * kXRowBytes); A(Xhdr, R * 256); A(xdirty, R * 8);
        if (cudaMemset(d.xdirty, 0xff, R * 8 * sizeof(unsigned)) != cudaSuccess) return cleanup(fail(FSB_ERR_CUDA, "cannot mark the packed rows"));
        int occx = 0;
        if (cudaFuncSetAttribute(fsb_passx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(XShared))) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occx, fsb_passx_kernel, kXThreads, sizeof(XShared)) != cudaSuccess || occx < 1)
            return cleanup(fail(FSB_ERR_CUDA, "fsb_passx_kernel cannot be resident"));
        if (const char *ev = getenv("FSB_PASS_CTAS")) occx = std::min(occx, std::max(1, atoi(ev)));
        e->grid_x = sms * occx;
    }
    A(beta, R * d.Mp); A(vol, R * d.Mp); A(impact, R * d.Mp);
    if (d.keep_history) A(volume, R * d.T * d.Mp);
    A(market, R * d.mkt_len);
    A(pos, R * d.N); A(horizon, R * d.N);
    A(stake, R * d.N); A(amount, R * d.N); A(threshold, R * d.N);
    A(nav_open, R * d.K); A(nav_close, R * d.K); A(cap0, R * d.K);
    A(hzero, R * d.K); A(stakeless, R * d.K); A(live_row, R * d.K);
    A(log_n, R);
    A(log_liquidated, R * d.cap_log); A(log_replacedby, R * d.cap_log); A(log_failtime, R * d.cap_log);
    A(log_ninv, R * d.cap_log); A(log_nstocks, R * d.cap_log);
    A(log_size, R * d.cap_log); A(log_liquidity, R * d.cap_log);
    A(err, R); A(cnt, R * kCnt);
    A(mkt_peak, R); A(mkt_dd, R);
    A(flow_hist, R * kFlowBins);
    d.scratch_ctas = std::max(e->grid[0], e->grid[2]);  // (only the event kernels spill order rows)
    if (d.big) d.scratch_ctas = std::max(d.scratch_ctas, e->grid[1]);  // (the pass kernel stages its columns in bigbuf)
    // big shapes run one batch at a time (an order-row scratch is cap_ev x Mp doubles per CTA: 11 MB at 2000 x 5000)
    const size_t scratch_slots = d.big ? 2 : 2 * kMaxBatches;
    A(scratch, scratch_slots * d.scratch_ctas * d.cap_ev * d.Mp);
    if (d.big) A(bigbuf, scratch_slots * d.scratch_ctas * static_cast<size_t>(d.big_stride));
    A(keys, R * d.kc * ((d.K + 3) & ~3));
    A(hand_i, R * d.hi_stride);
    A(hand_p, R * d.Mp);
    if (FSB_SPLIT_CHOICE && !d.big) A(hand_m, R * 2 * ((d.N + 63) / 64));
    if (d.keep_history) A(hand_s, R * d.Mp);
    A(queue, 8 * kMaxBatches);
#undef A
    {
        DevPoint *pd = nullptr; double *td = nullptr, *rt = nullptr;
        if (dalloc(e, &pd, e->points.size()) != FSB_OK || dalloc(e, &td, tables.size()) != FSB_OK || dalloc(e, &rt, d.K) != FSB_OK) return cleanup(FSB_ERR_CUDA);
        // rank probabilities of the rank rule: rank / normaliser in IEEE FP64, as functions.jl:405-406 divides
        std::vector<double> ranktab(d.K);
        const double normaliser = static_cast<double>(static_cast<long long>(d.K) * (d.K + 1) / 2);
        for (int i = 0; i < d.K; ++i) ranktab[i] = static_cast<double>(i + 1) / normaliser;
        if (cudaMemcpy(pd, e->points.data(), e->points.size() * sizeof(DevPoint), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(td, tables.data(), tables.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(rt, ranktab.data(), ranktab.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
            return cleanup(fail(FSB_ERR_CUDA, "cannot upload parameter tables"));
        d.points = pd; d.tables = td; d.ranktab = rt;
    }
    {   // per-replication pointer table of the step context (order: the pointer block of Ctx, fsb_step.cuh)
        const unsigned long long K = d.K, N = d.N, Mp = d.Mp, CL = d.cap_log;
        const unsigned long long kq = (d.K + 3) & ~3;
        struct { const void *base; unsigned long long stride; } rows[kRepPtrs] = {
            {d.H, static_cast<unsigned long long>(d.h_stride) * 8}, {d.P, static_cast<unsigned long long>(d.p_stride) * 8},
            {d.beta, Mp * 8}, {d.vol, Mp * 8}, {d.impact, Mp * 8},
            {d.volume, d.volume ? static_cast<unsigned long long>(d.T) * Mp * 8 : 0}, {d.market, static_cast<unsigned long long>(d.mkt_len) * 8},
            {d.stake, N * 8}, {d.amount, N * 8}, {d.threshold, N * 8}, {d.nav_open, K * 8}, {d.nav_close, K * 8},
            {d.pos, N * 4}, {d.horizon, N * 4}, {d.live_row, K * 4}, {d.cnt, static_cast<unsigned long long>(kCnt) * 4},
            {d.hzero, K}, {d.stakeless, K}, {d.mkt_peak, 8}, {d.mkt_dd, 8},
            {d.flow_hist, static_cast<unsigned long long>(kFlowBins) * 4}, {d.err, 4},
            {d.log_n, 4}, {d.log_liquidated, CL * 4}, {d.log_replacedby, CL * 4}, {d.log_failtime, CL * 4}, {d.log_ninv, CL * 4},
            {d.log_nstocks, CL * 4}, {d.log_size, CL * 8}, {d.log_liquidity, CL * 8},
            {d.keys, static_cast<unsigned long long>(d.kc) * kq * 8}};
        std::vector<ulonglong2> tab(kRepPtrs);
        for (int i = 0; i < kRepPtrs; ++i) tab[i] = make_ulonglong2(reinterpret_cast<unsigned long long>(rows[i].base), rows[i].stride);
        ulonglong2 *pt = nullptr;
        if (dalloc(e, &pt, static_cast<size_t>(kRepPtrs)) != FSB_OK) return cleanup(FSB_ERR_CUDA);
        if (cudaMemcpy(pt, tab.data(), tab.size() * sizeof(ulonglong2), cudaMemcpyHostToDevice) != cudaSuccess)
            return cleanup(fail(FSB_ERR_CUDA, "cannot upload the pointer table"));
        d.ptr_table = pt;
    }
    e->words_per_point = FSB_STATS_HEADER + (d.T + 1) + FSB_STATS_LIQ_BINS + FSB_STATS_DD_BINS + FSB_STATS_FLOW_BINS;
    if (dalloc(e, &e->stats_dev, static_cast<size_t>(n_points) * e->words_per_point) != FSB_OK) return cleanup(FSB_ERR_CUDA);
    e->z_mkt.assign(R, nullptr); e->z_stock.assign(R, nullptr); e->u_review.assign(R, nullptr);
    e->ev_tape.assign(R, nullptr); e->ev_n.assign(R, 0); e->trace_slot.assign(R, -1);
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->fork, cudaEventDisableTiming) != cudaSuccess)
        return cleanup(fail(FSB_ERR_CUDA, "cannot create stream/events"));
    for (int b = 0; b < kMaxBatches; ++b) {
        cudaStream_t st = nullptr; cudaEvent_t ev = nullptr;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
            return cleanup(fail(FSB_ERR_CUDA, "cannot create batch streams"));
        e->bstream.push_back(st); e->bjoin.push_back(ev);
    }
    *out = e;
    return FSB_OK;
}

extern "C" int32_t fsb_destroy(fsb_engine *e)
{
    if (!e) return FSB_OK;
    cudaSetDevice(e->cfg.device);
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (void *p : e->allocs) cudaFree(p);
    for (cudaEvent_t ev : e->kev) cudaEventDestroy(ev);
    for (cudaStream_t st : e->bstream) cudaStreamDestroy(st);
    for (cudaEvent_t ev : e->bjoin) cudaEventDestroy(ev);
    if (e->fork) cudaEventDestroy(e->fork);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    delete e;
    return FSB_OK;
}

static int32_t check_rep(fsb_engine *e, int64_t rep)
{
    if (!e) return fail(FSB_ERR_BAD_PARAMS, "null engine");
    if (rep < 0 || rep >= e->d.R) return fail(FSB_ERR_RANGE, "replication %lld out of range [0,%lld)", (long long)rep, e->d.R);
    CU(cudaSetDevice(e->cfg.device));
    // The engine's streams are non-blocking: the blocking copies of the upload / download / tape entry points (legacy
    // default stream) do not wait for them.  Every such entry point starts here, so work enqueued by fsb_run_async has
    // finished before state is read or overwritten (the batch streams are joined into e->stream by run_enqueue).
    CU(cudaStreamSynchronize(e->stream));
    return FSB_OK;
}

// ---- device initialisation ------------------------------------------------------------------
extern "C" int32_t fsb_init_device(fsb_engine *e, uint64_t seed)
{
    if (!e) return fail(FSB_ERR_BAD_PARAMS, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    e->d.seed = seed;
    fsb_init_kernel<<<static_cast<unsigned>(e->d.R), 256, e->init_smem, e->stream>>>(e->d);
    CU(cudaGetLastError());
    if (e->d.xdirty) CU(cudaMemsetAsync(e->d.xdirty, 0xff, static_cast<size_t>(e->d.R) * 8 * sizeof(unsigned), e->stream));  // every packed row is stale
    CU(cudaStreamSynchronize(e->stream));
    e->log_built = false;
    e->state_set = true;
    e->t_done = 0;
    return FSB_OK;
}

extern "C" int32_t fsb_set_seed(fsb_engine *e, uint64_t seed)
{
    if (!e) return fail(FSB_ERR_BAD_PARAMS, "null engine");
    e->d.seed = seed;
    return FSB_OK;
}

// ---- upload of the reference's structs (column-major, src/types.jl:15-50) ---------------------
extern "C" int32_t fsb_upload_state(fsb_engine *e, int64_t rep, const double *market, const double *stock_value,
                                    const double *beta, const double *vol, const double *impact, const double *volume,
                                    const double *inv_assets, const int64_t *horizon, const double *threshold,
                                    const double *holdings, const double *stakes, const double *fund_value)
{
    TRY(check_rep(e, rep));
    if (!market || !stock_value || !beta || !vol || !impact || !inv_assets || !horizon || !threshold || !holdings || !stakes)
        return fail(FSB_ERR_BAD_PARAMS, "fsb_upload_state: null array");
    const DevState &d = e->d;
    const int K = d.K, M = d.M, N = d.N, T = d.T, Mp = d.Mp, W = d.W;
    // investors: exactly one position per investor (SURVEY.md 3.2 invariant)
    std::vector<int> pos(N), hz(N);
    std::vector<double> amount(N), stake(N), thr(threshold, threshold + N);
    for (int n = 0; n < N; ++n) {
        int where = -1, count = 0;
        for (int c = 0; c <= K; ++c) {
            const double a = inv_assets[n + static_cast<size_t>(N) * c];
            if (a != 0.0) { where = c; ++count; }
        }
        if (count > 1) return fail(FSB_ERR_BAD_STATE, "investor %d holds %d positions; the engine needs at most one", n + 1, count);
        pos[n] = (where < 0) ? K : where;
        amount[n] = (where < 0) ? 0.0 : inv_assets[n + static_cast<size_t>(N) * where];
        if (amount[n] < 0.0) return fail(FSB_ERR_BAD_STATE, "investor %d has a negative position", n + 1);
        stake[n] = 0.0;
        for (int k = 0; k < K; ++k) {
            const double sv = stakes[k + static_cast<size_t>(K) * n];
            if (sv != 0.0) {
                if (k != pos[n]) return fail(FSB_ERR_BAD_STATE, "investor %d has a stake in fund %d but its position is in %d", n + 1, k + 1, pos[n] + 1);
                stake[n] = sv;
            }
        }
        if (horizon[n] < 1 || horizon[n] > W) return fail(FSB_ERR_BAD_STATE, "investor %d: horizon %lld outside 1..%d", n + 1, (long long)horizon[n], W);
        hz[n] = static_cast<int>(horizon[n]);
    }
    // holdings -> [K][Mp] asset-fastest
    std::vector<double> H(static_cast<size_t>(d.h_stride), 0.0);
    std::vector<unsigned char> hzero(K);
    std::vector<double> cap0(K);
    for (int k = 0; k < K; ++k) {
        bool nz = false;
        for (int m = 0; m < M; ++m) {
            const double h = holdings[k + static_cast<size_t>(K) * m];
            H[static_cast<size_t>(k) * Mp + m] = h;
            nz = nz || (h != 0.0);
        }
        hzero[k] = nz ? 0 : 1;
        if (fund_value) cap0[k] = fund_value[k];
        else {
            double c = 0.0;
            for (int n = 0; n < N; ++n) c += inv_assets[n + static_cast<size_t>(N) * k];
            cap0[k] = c;
        }
    }
    // prices: columns are contiguous in Julia's M x T layout; market value != 0 marks a produced column
    std::vector<double> P(static_cast<size_t>(d.p_stride), 0.0), mk(d.mkt_len, 0.0), volm;
    int64_t t_last = 0;
    for (int t = 1; t <= T; ++t) {
        const bool produced = (t <= W + 1) || market[t - 1] != 0.0;
        if (!produced) continue;
        t_last = t;
        for (int m = 0; m < M; ++m) P[static_cast<size_t>(pcol(d, t)) * Mp + m] = stock_value[m + static_cast<size_t>(M) * (t - 1)];
        mk[mslot(d, t)] = market[t - 1];
    }
    if (d.keep_history && volume) {
        volm.assign(static_cast<size_t>(T) * Mp, 0.0);
        for (int t = 1; t <= T; ++t)
            for (int m = 0; m < M; ++m) volm[static_cast<size_t>(t - 1) * Mp + m] = volume[m + static_cast<size_t>(M) * (t - 1)];
    }
    std::vector<double> b(Mp, 0.0), v(Mp, 0.0), im(Mp, 0.0);
    std::copy(beta, beta + M, b.begin()); std::copy(vol, vol + M, v.begin()); std::copy(impact, impact + M, im.begin());
    std::vector<int> live(K);
    for (int k = 0; k < K; ++k) live[k] = k;
    // respawn trigger of functions.jl:471: sum(stakes[k,:]) == 0, summed in ascending investor order
    std::vector<unsigned char> stakeless(K);
    for (int k = 0; k < K; ++k) {
        double c = 0.0;
        for (int n = 0; n < N; ++n)
            if (pos[n] == k) c = c + stake[n];
        stakeless[k] = (c == 0.0) ? 1 : 0;
    }
    double peak = 0.0;
    for (int t = 1; t <= T; ++t) peak = std::max(peak, market[t - 1]);
    const double zero = 0.0;
    const int izero = 0;
    std::vector<int> zeros_cnt(kCnt + kFlowBins, 0);
#define UP(dst, src, count) CU(cudaMemcpy((dst), (src), (count), cudaMemcpyHostToDevice))
    UP(d.H + rep * d.h_stride, H.data(), H.size() * 8);
    UP(d.P + rep * d.p_stride, P.data(), P.size() * 8);
    UP(d.beta + rep * Mp, b.data(), Mp * 8); UP(d.vol + rep * Mp, v.data(), Mp * 8); UP(d.impact + rep * Mp, im.data(), Mp * 8);
    if (d.keep_history && volume) UP(d.volume + rep * static_cast<long long>(T) * Mp, volm.data(), volm.size() * 8);
    UP(d.market + rep * d.mkt_len, mk.data(), mk.size() * 8);
    UP(d.pos + rep * N, pos.data(), N * 4); UP(d.horizon + rep * N, hz.data(), N * 4);
    UP(d.stake + rep * N, stake.data(), N * 8); UP(d.amount + rep * N, amount.data(), N * 8); UP(d.threshold + rep * N, thr.data(), N * 8);
    UP(d.cap0 + rep * K, cap0.data(), K * 8);
    UP(d.hzero + rep * K, hzero.data(), K); UP(d.stakeless + rep * K, stakeless.data(), K); UP(d.live_row + rep * K, live.data(), K * 4);
    UP(d.mkt_peak + rep, &mk[mslot(d, static_cast<int>(std::max<int64_t>(t_last, 1)))], 8);
    UP(d.mkt_dd + rep, &zero, 8);
    UP(d.err + rep, &izero, 4); UP(d.log_n + rep, &izero, 4);
    UP(d.cnt + rep * kCnt, zeros_cnt.data(), kCnt * 4);
    UP(d.flow_hist + rep * kFlowBins, zeros_cnt.data(), kFlowBins * 4);
#undef UP
    CU(cudaMemset(d.nav_open + rep * K, 0, K * 8));
    CU(cudaMemset(d.nav_close + rep * K, 0, K * 8));
    if (d.xdirty) CU(cudaMemset(d.xdirty + rep * 8, 0xff, 8 * sizeof(unsigned)));  // every packed row of this replication is stale
    (void)peak;
    e->log_built = false;
    e->state_set = true;
    e->t_done = (t_last > d.W + 1) ? t_last : 0;  // (a state uploaded mid-run resumes after its last produced column)
    return FSB_OK;
}

// ---- tapes / trace ------------------------------------------------------------------------------
static int32_t push_tables(fsb_engine *e)
{
    DevState &d = e->d;
    const size_t R = static_cast<size_t>(d.R);
    if (!e->z_mkt_d) {
        TRY(dalloc(e, &e->z_mkt_d, R)); TRY(dalloc(e, &e->z_stock_d, R)); TRY(dalloc(e, &e->u_review_d, R));
        TRY(dalloc(e, &e->ev_tape_d, R)); TRY(dalloc(e, &e->ev_n_d, R));
    }
    CU(cudaMemcpy(e->z_mkt_d, e->z_mkt.data(), R * sizeof(void *), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->z_stock_d, e->z_stock.data(), R * sizeof(void *), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->u_review_d, e->u_review.data(), R * sizeof(void *), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->ev_tape_d, e->ev_tape.data(), R * sizeof(void *), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->ev_n_d, e->ev_n.data(), R * sizeof(long long), cudaMemcpyHostToDevice));
    d.z_mkt = e->z_mkt_d; d.z_stock = e->z_stock_d; d.u_review = e->u_review_d;
    d.ev_tape = e->ev_tape_d; d.ev_n = e->ev_n_d;
    d.any_tape = 1;
    return FSB_OK;
}

extern "C" int32_t fsb_set_dense_tape(fsb_engine *e, int64_t rep, const double *z_mkt, const double *z_stock, const double *u_review)
{
    TRY(check_rep(e, rep));
    const DevState &d = e->d;
    auto up = [&](const double *src, size_t count, const double **slot) -> int32_t {
        if (!src) { *slot = nullptr; return FSB_OK; }
        double *p = nullptr;
        TRY(dalloc(e, &p, count, false));
        CU(cudaMemcpy(p, src, count * 8, cudaMemcpyHostToDevice));
        *slot = p;
        return FSB_OK;
    };
    TRY(up(z_mkt, d.T, &e->z_mkt[rep]));
    TRY(up(z_stock, static_cast<size_t>(d.M) * d.T, &e->z_stock[rep]));
    TRY(up(u_review, static_cast<size_t>(d.N) * d.T, &e->u_review[rep]));
    return push_tables(e);
}

extern "C" int32_t fsb_set_event_tape(fsb_engine *e, int64_t rep, const fsb_tape_rec *recs, int64_t n)
{
    TRY(check_rep(e, rep));
    if (n < 0 || (n > 0 && !recs)) return fail(FSB_ERR_BAD_PARAMS, "bad event tape");
    for (int64_t i = 1; i < n; ++i)
        if (recs[i - 1].key >= recs[i].key) return fail(FSB_ERR_BAD_PARAMS, "event tape must be strictly sorted by key");
    fsb_tape_rec *p = nullptr;
    TRY(dalloc(e, &p, static_cast<size_t>(std::max<int64_t>(n, 1)), false));
    if (n) CU(cudaMemcpy(p, recs, n * sizeof(fsb_tape_rec), cudaMemcpyHostToDevice));
    e->ev_tape[rep] = p;  // non-null pointer == tape mode for this replication, even when empty
    e->ev_n[rep] = n;
    return push_tables(e);
}

extern "C" int32_t fsb_set_trace(fsb_engine *e, int64_t rep, int32_t on)
{
    TRY(check_rep(e, rep));
    DevState &d = e->d;
    if (!on) {
        e->trace_slot[rep] = -1;
    } else if (e->trace_slot[rep] < 0) {
        // grow the slot arrays by one (rare, host-driven)
        const int ns = e->n_trace_slots + 1;
        double *nav = nullptr, *ps = nullptr, *pr = nullptr; fsb_trace_rec *recs = nullptr; int *tn = nullptr;
        const size_t kt = static_cast<size_t>(d.K) * d.T, mt = static_cast<size_t>(d.M) * d.T;
        TRY(dalloc(e, &nav, ns * kt)); TRY(dalloc(e, &ps, ns * mt)); TRY(dalloc(e, &pr, ns * mt));
        TRY(dalloc(e, &recs, static_cast<size_t>(ns) * d.cap_trace)); TRY(dalloc(e, &tn, ns));
        if (e->n_trace_slots) {
            const int os = e->n_trace_slots;
            CU(cudaMemcpy(nav, d.tr_nav, os * kt * 8, cudaMemcpyDeviceToDevice));
            CU(cudaMemcpy(ps, d.tr_psell, os * mt * 8, cudaMemcpyDeviceToDevice));
            CU(cudaMemcpy(pr, d.tr_presp, os * mt * 8, cudaMemcpyDeviceToDevice));
            CU(cudaMemcpy(recs, d.tr_recs, static_cast<size_t>(os) * d.cap_trace * sizeof(fsb_trace_rec), cudaMemcpyDeviceToDevice));
            CU(cudaMemcpy(tn, d.tr_n, os * 4, cudaMemcpyDeviceToDevice));
        }
        d.tr_nav = nav; d.tr_psell = ps; d.tr_presp = pr; d.tr_recs = recs; d.tr_n = tn;
        e->trace_slot[rep] = e->n_trace_slots;
        e->n_trace_slots = ns;
    }
    if (!e->trace_slot_d) TRY(dalloc(e, &e->trace_slot_d, static_cast<size_t>(d.R)));
    CU(cudaMemcpy(e->trace_slot_d, e->trace_slot.data(), d.R * 4, cudaMemcpyHostToDevice));
    d.trace_slot = e->trace_slot_d;
    return FSB_OK;
}

// ---- the hot path -----------------------------------------------------------------------------
// host buffers of a batched one-step tape (fsb_run_step_hosttape): copied slice by slice on the batch streams so
// that the transfers of one replication batch overlap the kernels of another
struct HostTape {
    const double *z_mkt = nullptr, *z_stock = nullptr, *u_review = nullptr;
    double *nav_open_out = nullptr;
};

static int32_t run_enqueue(fsb_engine *e, int64_t t_first, int64_t t_last, int32_t selector, int32_t flags,
                           int nb_override = 0, const HostTape *ht = nullptr)
{
    if (!e) return fail(FSB_ERR_BAD_PARAMS, "null engine");
    const DevState &d = e->d;
    if (selector < 0 || selector > 3) return fail(FSB_ERR_BAD_SELECTOR, "Please specify a valid fund selection method (best, goodenough, probabilistic, random)");
    if (!e->state_set) return fail(FSB_ERR_BAD_STATE, "no state: call fsb_upload_state or fsb_init_device first");
    if (t_first < d.W + 2 || t_last > d.T) return fail(FSB_ERR_RANGE, "time range [%lld,%lld] outside [%d,%d]", (long long)t_first, (long long)t_last, d.W + 2, d.T);
    {   // days are simulated once and in order: the price ring, the market slots and the log all assume it
        const int64_t t_next = e->t_done > 0 ? e->t_done + 1 : d.W + 2;
        if (t_last >= t_first && t_first != t_next)
            return fail(FSB_ERR_RANGE, "the next day to simulate is %lld, not %lld (re-running or skipping days would corrupt the price ring)", (long long)t_next, (long long)t_first);
    }
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaEventRecord(e->ev0, e->stream));
    e->last_launches = 0;
    e->kev_used = 0;
    if (!e->log_built) {
        fsb_build_log_kernel<<<static_cast<unsigned>(d.R), 256, 0, e->stream>>>(d);
        CU(cudaGetLastError());
        e->log_built = true;
        ++e->last_launches;
    }
    if (t_last >= t_first) {
        CU(cudaMemsetAsync(d.queue, 0, 8 * kMaxBatches * sizeof(int), e->stream));
        // one simulated day = five kernels per replication batch (draws, events-1, holdings pass, choice, events-2), every
        // batch on its own stream; the engine's stream forks before the first day and joins after the last
        int nb = nb_override > 0 ? nb_override : e->n_batches;
        if (d.big) nb = 1;
        nb = static_cast<int>(std::min<long long>(std::min<size_t>(nb, e->bstream.size()), d.R));
        if (nb < 1) nb = 1;
        if (nb > 1) {
            CU(cudaEventRecord(e->fork, e->stream));
            for (int b = 0; b < nb; ++b) CU(cudaStreamWaitEvent(e->bstream[b], e->fork, 0));
        }
        for (int64_t t = t_first; t <= t_last; ++t) {
            for (int b = 0; b < nb; ++b) {
                cudaStream_t st = nb > 1 ? e->bstream[b] : e->stream;
                StepArgs a{static_cast<int>(t), 1, selector, flags, b, (d.R * b) / nb, (d.R * (b + 1)) / nb};
                const long long rb = a.r_end - a.r_begin;
                if (ht) {  // this batch's slice of the step's shocks: host -> device on the batch's stream
                    CU(cudaMemcpyAsync(e->st_zmkt + a.r_begin, ht->z_mkt + a.r_begin, rb * 8, cudaMemcpyHostToDevice, st));
                    CU(cudaMemcpyAsync(e->st_zstock + a.r_begin * d.M, ht->z_stock + a.r_begin * d.M, rb * d.M * 8, cudaMemcpyHostToDevice, st));
                    CU(cudaMemcpyAsync(e->st_urev + a.r_begin * d.N, ht->u_review + a.r_begin * d.N, rb * d.N * 8, cudaMemcpyHostToDevice, st));
                }
                const unsigned g1 = static_cast<unsigned>(std::min<long long>(rb, e->grid[0])), g2 = static_cast<unsigned>(std::min<long long>(rb * d.pass_parts * d.pass_chunks, e->grid[1])),
                               g3 = static_cast<unsigned>(std::min<long long>(rb, e->grid[2]));
#if FSB_SPLIT_CHOICE
                if (!d.big) fsb_draws_kernel<<<static_cast<unsigned>(std::min<long long>((rb + 7) / 8, static_cast<long long>(e->sms) * 8)), 256, 0, st>>>(d, a);
#endif
                if (d.big) fsb_big_events1_kernel<<<g1, e->block, e->smem_bytes[0], st>>>(d, a, e->launch_parity);
                else if (e->fixed) fsb_fix_events1_kernel<<<g1, e->block, e->smem_bytes[0], st>>>(d, a, e->launch_parity);
                else fsb_events1_kernel<<<g1, e->block, e->smem_bytes[0], st>>>(d, a, e->launch_parity);
                const bool timed = e->kev_used + 2 <= 4096;  // (the dominant kernel is timed launch by launch: bench.py's roofline)
                if (timed) {
                    while (e->kev.size() < e->kev_used + 2) { cudaEvent_t ev = nullptr; CU(cudaEventCreate(&ev)); e->kev.push_back(ev); }
                    CU(cudaEventRecord(e->kev[e->kev_used], st));
                }
                if (d.big) fsb_big_pass_kernel<<<g2, e->block_pass, e->smem_bytes[1], st>>>(d, a, e->launch_parity, e->tm_h, e->tm_h16);
                else if (e->usex) fsb_passx_kernel<<<static_cast<unsigned>(std::min<long long>(rb, e->grid_x)), kXThreads, sizeof(XShared), st>>>(d, a, e->launch_parity);
                else fsb_pass_kernel<<<g2, e->block_pass, e->smem_bytes[1], st>>>(d, a, e->launch_parity, e->tm_h, e->tm_h16);
                if (timed) { CU(cudaEventRecord(e->kev[e->kev_used + 1], st)); e->kev_used += 2; }
                if (ht && ht->nav_open_out)  // the NAVs of :460 are final after the pass: device -> host while events-2 runs
                    CU(cudaMemcpyAsync(ht->nav_open_out + a.r_begin * d.K, d.nav_open + a.r_begin * d.K, rb * d.K * 8, cudaMemcpyDeviceToHost, st));
#if FSB_SPLIT_CHOICE
                if (e->fixed) fsb_fix_choice_kernel<<<g3, e->block_e2, e->smem_bytes[2], st>>>(d, a, e->launch_parity);
                else if (!d.big) fsb_choice_kernel<<<g3, e->block_e2, e->smem_bytes[2], st>>>(d, a, e->launch_parity);
#endif
                if (d.big) fsb_big_events2_kernel<<<g3, e->block_e2, e->smem_bytes[2], st>>>(d, a, e->launch_parity);
                else if (e->fixed) fsb_fix_events2_kernel<<<g3, e->block_e2, e->smem_bytes[2], st>>>(d, a, e->launch_parity);
                else fsb_events2_kernel<<<g3, e->block_e2, e->smem_bytes[2], st>>>(d, a, e->launch_parity);
                e->last_launches += (FSB_SPLIT_CHOICE && !d.big) ? 5 : 3;
            }
            e->launch_parity ^= 1;
        }
        if (nb > 1)
            for (int b = 0; b < nb; ++b) {
                CU(cudaEventRecord(e->bjoin[b], e->bstream[b]));
                CU(cudaStreamWaitEvent(e->stream, e->bjoin[b], 0));
            }
        CU(cudaGetLastError());
        e->t_done = t_last;
    }
    CU(cudaEventRecord(e->ev1, e->stream));
    return FSB_OK;
}

extern "C" int32_t fsb_sync(fsb_engine *e)
{
    if (!e) return fail(FSB_ERR_BAD_PARAMS, "null engine");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) e->last_total_ms = ms;
    else (void)cudaGetLastError();
    e->last_pass_ms = 0.0;
    e->last_pass_launches = 0;
    for (size_t q = 0; q + 1 < e->kev_used; q += 2) {
        float pm = 0.f;
        if (cudaEventElapsedTime(&pm, e->kev[q], e->kev[q + 1]) == cudaSuccess) { e->last_pass_ms += pm; ++e->last_pass_launches; }
        else (void)cudaGetLastError();
    }
    return FSB_OK;
}

extern "C" int32_t fsb_run_async(fsb_engine *e, int64_t t_first, int64_t t_last, int32_t selector, int32_t flags)
{
    return run_enqueue(e, t_first, t_last, selector, flags);
}

extern "C" int32_t fsb_run(fsb_engine *e, int64_t t_first, int64_t t_last, int32_t selector, int32_t flags)
{
    TRY(run_enqueue(e, t_first, t_last, selector, flags));
    TRY(fsb_sync(e));
    // surface replication errors (capacity overflow, Q6, tape problems) as a status
    std::vector<int> err(static_cast<size_t>(e->d.R));
    CU(cudaMemcpy(err.data(), e->d.err, err.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < err.size(); ++r)
        if (err[r] & ~FSB_REP_TRACE_OVERFLOW)
            return fail(FSB_ERR_REPLICATION, "replication %zu raised error flags 0x%x (see fsb_rep_status)", r, err[r]);
    return FSB_OK;
}

extern "C" int32_t fsb_run_step_hosttape(fsb_engine *e, int64_t t, int32_t selector, int32_t flags, const double *z_mkt,
                                         const double *z_stock, const double *u_review, double *nav_open_out)
{
    if (!e || !z_mkt || !z_stock || !u_review) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    DevState &d = e->d;
    CU(cudaSetDevice(e->cfg.device));
    const size_t R = static_cast<size_t>(d.R);
    if (!e->st_zmkt) {
        TRY(dalloc(e, &e->st_zmkt, R, false)); TRY(dalloc(e, &e->st_zstock, R * d.M, false)); TRY(dalloc(e, &e->st_urev, R * d.N, false));
    }
    d.st_zmkt = e->st_zmkt; d.st_zstock = e->st_zstock; d.st_urev = e->st_urev;
    d.step_tape = 1;
    HostTape ht;
    ht.z_mkt = z_mkt; ht.z_stock = z_stock; ht.u_review = u_review; ht.nav_open_out = nav_open_out;
    int nbh = 8;  // (measured: 1 -> 4.9 ms per step of 10^4 replications, 4 -> 3.40, 8 -> 3.25, 12 -> 3.24: the host link is the limit then)
    if (const char *ev = getenv("FSB_HOSTTAPE_BATCHES")) nbh = std::max(1, std::min(kMaxBatches, atoi(ev)));
    const int32_t rc = run_enqueue(e, t, t, selector, flags, nbh, &ht);
    d.step_tape = 0;
    TRY(rc);
    return fsb_sync(e);
}

extern "C" int32_t fsb_last_run_timing(fsb_engine *e, double *total_ms, double *step_kernel_ms, int64_t *launches)
{
    if (!e) return fail(FSB_ERR_BAD_PARAMS, "null engine");
    if (total_ms) *total_ms = e->last_total_ms;
    if (step_kernel_ms) *step_kernel_ms = e->last_pass_ms;  // summed over the pass-kernel launches (the dominant kernel)
    if (launches) *launches = e->last_launches;
    return FSB_OK;
}

extern "C" int64_t fsb_last_pass_launches(const fsb_engine *e) { return e ? e->last_pass_launches : 0; }

extern "C" int32_t fsb_rep_status(fsb_engine *e, int32_t *flags)
{
    if (!e || !flags) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));  // (see check_rep)
    CU(cudaMemcpy(flags, e->d.err, e->d.R * 4, cudaMemcpyDeviceToHost));
    return FSB_OK;
}

// ---- downloads ----------------------------------------------------------------------------------
extern "C" int32_t fsb_download_compact(fsb_engine *e, int64_t rep, double *holdings, double *price_now, int64_t *inv_fund,
                                        double *inv_amount, double *inv_stake, double *nav_open, double *nav_close,
                                        double *market_now)
{
    TRY(check_rep(e, rep));
    const DevState &d = e->d;
    const int K = d.K, M = d.M, N = d.N, Mp = d.Mp;
    const int t_now = static_cast<int>(e->t_done > 0 ? e->t_done : d.W + 1);
    if (holdings) {
        std::vector<double> H(static_cast<size_t>(K) * Mp);
        CU(cudaMemcpy(H.data(), d.H + rep * d.h_stride, H.size() * 8, cudaMemcpyDeviceToHost));
        for (int k = 0; k < K; ++k)
            for (int m = 0; m < M; ++m) holdings[k + static_cast<size_t>(K) * m] = H[static_cast<size_t>(k) * Mp + m];
    }
    if (price_now) CU(cudaMemcpy(price_now, d.P + rep * d.p_stride + static_cast<long long>(pcol(d, t_now)) * Mp, M * 8, cudaMemcpyDeviceToHost));
    if (inv_fund) {
        std::vector<int> pos(N);
        CU(cudaMemcpy(pos.data(), d.pos + rep * N, N * 4, cudaMemcpyDeviceToHost));
        for (int n = 0; n < N; ++n) inv_fund[n] = pos[n] + 1;
    }
    if (inv_amount) CU(cudaMemcpy(inv_amount, d.amount + rep * N, N * 8, cudaMemcpyDeviceToHost));
    if (inv_stake) CU(cudaMemcpy(inv_stake, d.stake + rep * N, N * 8, cudaMemcpyDeviceToHost));
    if (nav_open) CU(cudaMemcpy(nav_open, d.nav_open + rep * K, K * 8, cudaMemcpyDeviceToHost));
    if (nav_close) CU(cudaMemcpy(nav_close, d.nav_close + rep * K, K * 8, cudaMemcpyDeviceToHost));
    if (market_now) CU(cudaMemcpy(market_now, d.market + rep * d.mkt_len + mslot(d, t_now), 8, cudaMemcpyDeviceToHost));
    return FSB_OK;
}

extern "C" int32_t fsb_download_state(fsb_engine *e, int64_t rep, double *market, double *stock_value, double *beta, double *vol,
                                      double *impact, double *volume, double *inv_assets, int64_t *horizon, double *threshold,
                                      double *holdings, double *stakes, double *fund_value)
{
    TRY(check_rep(e, rep));
    const DevState &d = e->d;
    if (!d.keep_history) return fail(FSB_ERR_NO_HISTORY, "fsb_download_state needs an engine created with keep_history=1");
    const int K = d.K, M = d.M, N = d.N, T = d.T, Mp = d.Mp;
    if (market) CU(cudaMemcpy(market, d.market + rep * d.mkt_len, T * 8, cudaMemcpyDeviceToHost));
    if (stock_value || volume) {
        std::vector<double> buf(static_cast<size_t>(T) * Mp);
        if (stock_value) {
            CU(cudaMemcpy(buf.data(), d.P + rep * d.p_stride, buf.size() * 8, cudaMemcpyDeviceToHost));
            for (int t = 0; t < T; ++t)
                for (int m = 0; m < M; ++m) stock_value[m + static_cast<size_t>(M) * t] = buf[static_cast<size_t>(t) * Mp + m];
        }
        if (volume) {
            CU(cudaMemcpy(buf.data(), d.volume + rep * static_cast<long long>(T) * Mp, buf.size() * 8, cudaMemcpyDeviceToHost));
            for (int t = 0; t < T; ++t)
                for (int m = 0; m < M; ++m) volume[m + static_cast<size_t>(M) * t] = buf[static_cast<size_t>(t) * Mp + m];
        }
    }
    if (beta) CU(cudaMemcpy(beta, d.beta + rep * Mp, M * 8, cudaMemcpyDeviceToHost));
    if (vol) CU(cudaMemcpy(vol, d.vol + rep * Mp, M * 8, cudaMemcpyDeviceToHost));
    if (impact) CU(cudaMemcpy(impact, d.impact + rep * Mp, M * 8, cudaMemcpyDeviceToHost));
    std::vector<int> pos(N), hz(N);
    std::vector<double> amount(N), stake(N);
    CU(cudaMemcpy(pos.data(), d.pos + rep * N, N * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hz.data(), d.horizon + rep * N, N * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(amount.data(), d.amount + rep * N, N * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(stake.data(), d.stake + rep * N, N * 8, cudaMemcpyDeviceToHost));
    if (inv_assets) {
        std::fill(inv_assets, inv_assets + static_cast<size_t>(N) * (K + 1), 0.0);
        for (int n = 0; n < N; ++n) inv_assets[n + static_cast<size_t>(N) * pos[n]] = amount[n];
    }
    if (horizon) for (int n = 0; n < N; ++n) horizon[n] = hz[n];
    if (threshold) CU(cudaMemcpy(threshold, d.threshold + rep * N, N * 8, cudaMemcpyDeviceToHost));
    if (stakes) {
        std::fill(stakes, stakes + static_cast<size_t>(K) * N, 0.0);
        for (int n = 0; n < N; ++n) if (pos[n] < K) stakes[pos[n] + static_cast<size_t>(K) * n] = stake[n];
    }
    if (holdings) TRY(fsb_download_compact(e, rep, holdings, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    if (fund_value) {
        double *out = nullptr;
        CU(cudaMalloc(&out, static_cast<size_t>(K) * T * 8));
        fsb_fund_value_kernel<<<148, 256, 0, e->stream>>>(d, rep, static_cast<int>(e->t_done), out);
        std::vector<double> buf(static_cast<size_t>(K) * T);
        cudaError_t ce = cudaStreamSynchronize(e->stream);
        if (ce == cudaSuccess) ce = cudaMemcpy(buf.data(), out, buf.size() * 8, cudaMemcpyDeviceToHost);
        cudaFree(out);
        if (ce != cudaSuccess) return fail(FSB_ERR_CUDA, "fund value download: %s", cudaGetErrorString(ce));
        for (int k = 0; k < K; ++k)
            for (int t = 0; t < T; ++t) fund_value[k + static_cast<size_t>(K) * t] = buf[static_cast<size_t>(k) * T + t];
        if (e->t_done == 0) {  // before any step the reference still holds the capital in column 1
            std::vector<double> c0(K);
            CU(cudaMemcpy(c0.data(), d.cap0 + rep * K, K * 8, cudaMemcpyDeviceToHost));
            for (int k = 0; k < K; ++k) fund_value[k] = c0[k];
        }
    }
    return FSB_OK;
}

extern "C" int32_t fsb_download_log(fsb_engine *e, int64_t rep, int64_t *n_rows, int64_t *liquidated, int64_t *replacedby,
                                    double *initialsize, double *liquidity, int64_t *failtime, int64_t *ninvestors,
                                    int64_t *nstocks)
{
    TRY(check_rep(e, rep));
    const DevState &d = e->d;
    if (!e->log_built) {  // the reference builds the log at the top of modelrun; do the same on demand
        if (!e->state_set) return fail(FSB_ERR_BAD_STATE, "no state");
        fsb_build_log_kernel<<<static_cast<unsigned>(d.R), 256, 0, e->stream>>>(d);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(e->stream));
        e->log_built = true;
    }
    int rows = 0;
    CU(cudaMemcpy(&rows, d.log_n + rep, 4, cudaMemcpyDeviceToHost));
    if (n_rows) *n_rows = rows;
    const long long lb = rep * d.cap_log;
    auto geti = [&](const int *src, int64_t *dst) -> int32_t {
        if (!dst) return FSB_OK;
        std::vector<int> b(rows);
        CU(cudaMemcpy(b.data(), src + lb, rows * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < rows; ++i) dst[i] = b[i];
        return FSB_OK;
    };
    TRY(geti(d.log_liquidated, liquidated)); TRY(geti(d.log_replacedby, replacedby)); TRY(geti(d.log_failtime, failtime));
    TRY(geti(d.log_ninv, ninvestors)); TRY(geti(d.log_nstocks, nstocks));
    if (initialsize) CU(cudaMemcpy(initialsize, d.log_size + lb, rows * 8, cudaMemcpyDeviceToHost));
    if (liquidity) CU(cudaMemcpy(liquidity, d.log_liquidity + lb, rows * 8, cudaMemcpyDeviceToHost));
    return FSB_OK;
}

extern "C" int32_t fsb_stylised_facts(fsb_engine *e, int64_t rep_first, int64_t n_reps, int32_t maxlag, double returnspctile,
                                      double *stock_pacf, double *market_pacf, double *stock_volacluster,
                                      double *market_volacluster, double *kurtoses, double *lossgain, double *corrs)
{
    if (!e) return fail(FSB_ERR_BAD_PARAMS, "null engine");
    if (n_reps < 1) return fail(FSB_ERR_RANGE, "n_reps %lld", (long long)n_reps);
    TRY(check_rep(e, rep_first));
    TRY(check_rep(e, rep_first + n_reps - 1));
    const DevState &d = e->d;
    if (!d.keep_history) return fail(FSB_ERR_NO_HISTORY, "fsb_stylised_facts needs an engine created with keep_history=1");
    if (!e->state_set) return fail(FSB_ERR_BAD_STATE, "no state");
    const int M = d.M, n = d.T - d.W - 1, n2 = n - d.W;
    if (maxlag < 1 || maxlag > facts_max_lag()) return fail(FSB_ERR_BAD_PARAMS, "maxlag %d not in 1..%d", maxlag, facts_max_lag());
    if (n <= 2 * maxlag + 1) return fail(FSB_ERR_BAD_PARAMS, "%d returns after the burn-in are too few for %d lags", n, maxlag);
    if (!(returnspctile >= 0.0 && returnspctile <= 100.0)) return fail(FSB_ERR_BAD_PARAMS, "returnspctile %g not in [0,100]", returnspctile);
    if (n2 >= 1 && static_cast<int>(static_cast<double>(n2 - 1) * (returnspctile / 100.0)) + 2 > facts_max_sel())
        return fail(FSB_ERR_BAD_PARAMS, "percentile %g of %d returns needs more than the %d order statistics a thread keeps",
                    returnspctile, n2, facts_max_sel());
    // The kernels write the seven arrays in the caller's layouts and each goes straight to the caller's memory.
    // Replications go in chunks through two device buffers: a copy into pageable host memory blocks the host, so
    // the kernels of chunk c+1 are launched (engine stream) before the copies of chunk c start (a batch stream).
    const size_t per_rep = static_cast<size_t>(facts_doubles_per_rep(M, maxlag));
    // 296 replications = 592 CTAs = one wave of 4 CTAs per SM; measured at 2048 replications (wall / sum of kernel
    // times, ms; before the batched heap insertions): 296: 31.3 / 25.6, 592: 33.3 / 23.9, 1024: 36.2 / 21.5, one chunk:
    // 45.3 / 19.9; with them 296: 29.2 / 17.6, one chunk: 41.1 / 15.5
    const char *ck = getenv("FSB_FACTS_CHUNK");  // measurement knob
    const int64_t chunk = std::min<int64_t>(n_reps, ck && atoll(ck) > 0 ? atoll(ck) : 296);
    const int64_t n_chunks = (n_reps + chunk - 1) / chunk;
    cudaStream_t cs = e->bstream.empty() ? e->stream : e->bstream[0];
    double *dev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev(static_cast<size_t>(2 * n_chunks), nullptr);
    cudaError_t ce = cudaMalloc(&dev[0], per_rep * chunk * 8);
    if (ce == cudaSuccess && n_chunks > 1) ce = cudaMalloc(&dev[1], per_rep * chunk * 8);
    for (size_t i = 0; i < ev.size() && ce == cudaSuccess; ++i) ce = cudaEventCreate(&ev[i]);
    auto launch = [&](int64_t c) {
        const int64_t c0 = c * chunk, nc = std::min(chunk, n_reps - c0);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[2 * c], e->stream);
        if (ce == cudaSuccess) ce = facts_launch(d, e->stream, rep_first + c0, static_cast<int>(nc), maxlag, returnspctile, dev[c & 1]);
        if (ce == cudaSuccess) ce = cudaEventRecord(ev[2 * c + 1], e->stream);
    };
    if (ce == cudaSuccess) launch(0);
    for (int64_t c = 0; c < n_chunks && ce == cudaSuccess; ++c) {
        if (c + 1 < n_chunks) launch(c + 1);  // its buffer was read by the copies of chunk c-1, which have returned
        const int64_t c0 = c * chunk, nc = std::min(chunk, n_reps - c0);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(cs, ev[2 * c + 1], 0);
        const double *src = dev[c & 1];
        auto out = [&](double *dst, size_t per) {  // per: doubles per replication of this array
            if (dst && ce == cudaSuccess)
                ce = cudaMemcpyAsync(dst + per * c0, src, per * nc * 8, cudaMemcpyDeviceToHost, cs);
            src += per * nc;
        };
        const size_t lm = static_cast<size_t>(maxlag) * M;
        out(stock_pacf, lm); out(stock_volacluster, lm); out(market_pacf, maxlag); out(market_volacluster, maxlag);
        out(kurtoses, M); out(lossgain, M); out(corrs, M);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(cs);  // (pinned destinations would make the copies asynchronous)
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    e->last_facts_ms = 0.0;
    for (int64_t c = 0; c < n_chunks && ce == cudaSuccess; ++c) {
        float ms = 0.f;
        ce = cudaEventElapsedTime(&ms, ev[2 * c], ev[2 * c + 1]);
        e->last_facts_ms += ms;
    }
    if (ce != cudaSuccess) cudaDeviceSynchronize();  // nothing may still use the buffers freed below
    for (cudaEvent_t x : ev) if (x) cudaEventDestroy(x);
    cudaFree(dev[0]);
    cudaFree(dev[1]);
    if (ce != cudaSuccess) return fail(FSB_ERR_CUDA, "stylised facts: %s", cudaGetErrorString(ce));
    return FSB_OK;
}

extern "C" int32_t fsb_last_facts_timing(fsb_engine *e, double *kernel_ms)
{
    if (!e || !kernel_ms) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    *kernel_ms = e->last_facts_ms;
    return FSB_OK;
}

extern "C" int32_t fsb_download_trace(fsb_engine *e, int64_t rep, double *nav_open, double *p_sell, double *p_resp,
                                      fsb_trace_rec *recs, int64_t *n_recs)
{
    TRY(check_rep(e, rep));
    const DevState &d = e->d;
    const int slot = e->trace_slot[rep];
    if (slot < 0) return fail(FSB_ERR_BAD_PARAMS, "replication %lld is not traced", (long long)rep);
    const size_t kt = static_cast<size_t>(d.K) * d.T, mt = static_cast<size_t>(d.M) * d.T;
    if (nav_open) CU(cudaMemcpy(nav_open, d.tr_nav + slot * kt, kt * 8, cudaMemcpyDeviceToHost));
    if (p_sell) CU(cudaMemcpy(p_sell, d.tr_psell + slot * mt, mt * 8, cudaMemcpyDeviceToHost));
    if (p_resp) CU(cudaMemcpy(p_resp, d.tr_presp + slot * mt, mt * 8, cudaMemcpyDeviceToHost));
    int n = 0;
    CU(cudaMemcpy(&n, d.tr_n + slot, 4, cudaMemcpyDeviceToHost));
    if (recs) CU(cudaMemcpy(recs, d.tr_recs + static_cast<size_t>(slot) * d.cap_trace, static_cast<size_t>(n) * sizeof(fsb_trace_rec), cudaMemcpyDeviceToHost));
    if (n_recs) *n_recs = n;
    return FSB_OK;
}

// ---- statistics + NCCL ---------------------------------------------------------------------------
extern "C" int32_t fsb_comm_unique_id(uint8_t id[128])
{
    TRY(nccl_load());
    ncclUniqueId uid;
    NC(g_nccl.GetUniqueId(&uid));
    static_assert(sizeof(uid) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &uid, 128);
    return FSB_OK;
}

extern "C" int32_t fsb_comm_init_rank(fsb_engine *e, int32_t n_ranks, int32_t rank, const uint8_t id[128])
{
    if (!e) return fail(FSB_ERR_BAD_PARAMS, "null engine");
    TRY(nccl_load());
    CU(cudaSetDevice(e->cfg.device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    NC(g_nccl.CommInitRank(&e->comm, n_ranks, uid, rank));
    e->n_ranks = n_ranks;
    return FSB_OK;
}

extern "C" int32_t fsb_comm_init_all(fsb_engine **engines, int32_t n)
{
    if (!engines || n < 1) return fail(FSB_ERR_BAD_PARAMS, "bad engine list");
    TRY(nccl_load());
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) devs[i] = engines[i]->cfg.device;
    NC(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) { engines[i]->comm = comms[i]; engines[i]->n_ranks = n; }
    return FSB_OK;
}

static int32_t stats_local(fsb_engine *e)
{
    const DevState &d = e->d;
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaMemsetAsync(e->stats_dev, 0, static_cast<size_t>(d.n_points) * e->words_per_point * 8, e->stream));
    if (!e->log_built) {
        fsb_build_log_kernel<<<static_cast<unsigned>(d.R), 256, 0, e->stream>>>(d);
        e->log_built = true;
    }
    fsb_stats_kernel<<<static_cast<unsigned>(d.R), 128, 0, e->stream>>>(d, e->stats_dev, e->words_per_point);
    CU(cudaGetLastError());
    return FSB_OK;
}

extern "C" int32_t fsb_stats(fsb_engine *e, uint64_t *out)
{
    if (!e || !out) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    TRY(stats_local(e));
    const size_t words = static_cast<size_t>(e->d.n_points) * e->words_per_point;
    if (e->comm)  // the single collective of the whole run
        NC(g_nccl.AllReduce(e->stats_dev, e->stats_dev, words, ncclUint64, ncclSum, e->comm, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(out, e->stats_dev, words * 8, cudaMemcpyDeviceToHost));
    return FSB_OK;
}

extern "C" int32_t fsb_stats_local(fsb_engine *e, uint64_t *out)
{
    if (!e || !out) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    TRY(stats_local(e));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaMemcpy(out, e->stats_dev, static_cast<size_t>(e->d.n_points) * e->words_per_point * 8, cudaMemcpyDeviceToHost));
    return FSB_OK;
}

extern "C" int32_t fsb_stats_all(fsb_engine **engines, int32_t n, uint64_t *out)
{
    if (!engines || n < 1 || !out) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    for (int i = 0; i < n; ++i) TRY(stats_local(engines[i]));
    const size_t words = static_cast<size_t>(engines[0]->d.n_points) * engines[0]->words_per_point;
    if (n > 1) {
        for (int i = 0; i < n; ++i) if (!engines[i]->comm) return fail(FSB_ERR_NCCL, "call fsb_comm_init_all first");
        NC(g_nccl.GroupStart());
        for (int i = 0; i < n; ++i) {
            CU(cudaSetDevice(engines[i]->cfg.device));
            NC(g_nccl.AllReduce(engines[i]->stats_dev, engines[i]->stats_dev, words, ncclUint64, ncclSum, engines[i]->comm, engines[i]->stream));
        }
        NC(g_nccl.GroupEnd());
    }
    for (int i = 0; i < n; ++i) {
        CU(cudaSetDevice(engines[i]->cfg.device));
        CU(cudaStreamSynchronize(engines[i]->stream));
    }
    CU(cudaSetDevice(engines[0]->cfg.device));
    CU(cudaMemcpy(out, engines[0]->stats_dev, words * 8, cudaMemcpyDeviceToHost));
    return FSB_OK;
}

// ---- debug --------------------------------------------------------------------------------------
// FSB_REDZONE=1 engines: number of device buffers whose guard bands (256 bytes before, >= 256 after) no longer hold
// the fill pattern, i.e. that some kernel wrote outside of; -1 in *n_buffers when the engine was created without red zones
// the column chunks the big-shape pass makes of a day with ncols distinct horizons (fsb_step.cuh: big_chunk_count / big_chunk,
// the functions the kernel itself calls): first[i], count[i] for chunk i; returns the number of chunks (<= cap, else -1)
extern "C" int32_t fsb_debug_pass_chunks(int32_t ncols, int32_t *first, int32_t *count, int32_t cap)
{
    if (ncols < 0 || !first || !count) return -1;
    const int nch = big_chunk_count(ncols);
    if (nch > cap) return -1;
    for (int ch = 0; ch < nch; ++ch) {
        int c0 = 0, nc = 0;
        big_chunk(ncols, nch, ch, &c0, &nc);
        first[ch] = c0; count[ch] = nc;
    }
    return nch;
}

extern "C" int32_t fsb_debug_redzones(fsb_engine *e, int64_t *n_buffers, int64_t *n_damaged)
{
    if (!e || !n_buffers || !n_damaged) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    *n_buffers = e->redzone ? static_cast<int64_t>(e->zones.size()) : -1;
    *n_damaged = 0;
    std::vector<unsigned char> band(512);
    for (const auto &z : e->zones) {
        const size_t padded = (z.second + 255) & ~static_cast<size_t>(255);
        bool bad = false;
        CU(cudaMemcpy(band.data(), z.first - 256, 256, cudaMemcpyDeviceToHost));
        for (int i = 0; i < 256; ++i) bad = bad || band[i] != 0xA5;
        const size_t tail = padded - z.second + 256;  // the slack up to the 256-byte boundary belongs to the band
        CU(cudaMemcpy(band.data(), z.first + z.second, tail, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tail; ++i) bad = bad || band[i] != 0xA5;
        *n_damaged += bad ? 1 : 0;
    }
    return FSB_OK;
}

// per-phase cycle counters of the step kernel (all zero unless built with -DFSB_PHASE_TIMING); reset = 1 clears them
extern "C" int32_t fsb_debug_phase_cycles(fsb_engine *e, uint64_t out[32], int32_t reset)
{
    if (!e || !out) return fail(FSB_ERR_BAD_PARAMS, "null argument");
    CU(cudaSetDevice(e->cfg.device));
    CU(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < 32; ++i) out[i] = 0;
#ifdef FSB_PHASE_TIMING
    CU(cudaMemcpyFromSymbol(out, g_phase_cycles, 32 * sizeof(unsigned long long)));
    if (reset) {
        unsigned long long z[32] = {0};
        CU(cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z)));
    }
#else
    (void)reset;
#endif
    return FSB_OK;
}

extern "C" int32_t fsb_debug_draws(fsb_engine *e, uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b,
                                   uint32_t idx0, int32_t n, int32_t kind, uint32_t n_discrete, double *out)
{
    if (!e || !out || n < 1) return fail(FSB_ERR_BAD_PARAMS, "bad argument");
    CU(cudaSetDevice(e->cfg.device));
    double *dv = nullptr;
    CU(cudaMalloc(&dv, static_cast<size_t>(n) * 8));
    fsb_debug_draws_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(seed, rep, site, a, b, idx0, n, kind, n_discrete, dv);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpy(out, dv, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost);
    cudaFree(dv);
    if (ce != cudaSuccess) return fail(FSB_ERR_CUDA, "debug draws: %s", cudaGetErrorString(ce));
    return FSB_OK;
}

extern "C" int32_t fsb_debug_dot(fsb_engine *e, const double *x, const double *y, int64_t n, double *out)
{
    if (!e || !x || !y || !out || n < 1) return fail(FSB_ERR_BAD_PARAMS, "bad argument");
    CU(cudaSetDevice(e->cfg.device));
    const size_t np = static_cast<size_t>((n + 1) & ~1LL);
    double *dx = nullptr, *dy = nullptr, *dr = nullptr;
    CU(cudaMalloc(&dx, np * 8)); CU(cudaMalloc(&dy, np * 8)); CU(cudaMalloc(&dr, 8));
    cudaMemset(dx, 0, np * 8); cudaMemset(dy, 0, np * 8);
    cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dy, y, n * 8, cudaMemcpyHostToDevice);
    fsb_debug_dot_kernel<<<1, 32, 0, e->stream>>>(dx, dy, static_cast<int>(n), dr);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpy(out, dr, 8, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dy); cudaFree(dr);
    if (ce != cudaSuccess) return fail(FSB_ERR_CUDA, "debug dot: %s", cudaGetErrorString(ce));
    return FSB_OK;
}
