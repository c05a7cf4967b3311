// fsb_passx.cuh -- the holdings pass of the default shape (Mp == 256, K <= 256) over a PACKED MIRROR of the holdings.
//
// fundholdinit! picks at most portfsizerange[end] stocks per fund (functions.jl:119-133), liquidate! scales a
// row (:194), reinvest! buys in proportion to the current weights (:367-368): a row's sparsity pattern only
// changes at respawn (:291-292).  The dense K x M holdings (fsb_state.h: H) stay the authoritative state that
// every event kernel reads and writes; next to them the engine keeps a packed copy X of every row, and the
// pass -- the one kernel that has to stream ALL rows every day -- reads the copy: ~0.7 KB instead of 2 KB
// per row at the default Params, and 1/4 of the multiply-adds.
//
// Bit-exactness.  The canonical row.column product is 8 chains: chain j owns the elements
// 16i+2j, 16i+2j+1 (ascending) and accumulates acc = fma(h, p, acc) from +0.0; the chains are combined by the
// butterfly 4, 2, 1.  A zero holding contributes fma(0, p, acc) == acc exactly for every FINITE p (acc is never
// -0.0: it starts as +0.0 and x + (-x) rounds to +0.0), so a chain may skip its zeros.  A packed row therefore
// stores, per chain j, the chain's non-zero elements in ascending order, padded with (stock 2j, 0.0) to the
// length L of the row's longest chain rounded up to even; slots come in pairs: pair p = (it >> 1) of chain j sits at
// position 8 * p + j, so a lane reads its two slots with one 16-byte load and the 8 lanes of a row read 128 contiguous
// bytes per pair.  Layout of a row (kXRowBytes):
//     [0, 128)      stock codes, one byte each: byte 2 * (8 * p + j) + (it & 1) holds e = 2 * (m >> 4) + (m & 1), the stock's
//                   position in its chain (the chain is the lane's own: m = 16 * (e >> 1) + 2 * j + (e & 1))
//     [128, 1152)   values:  double 2 * (8 * p + j) + (it & 1)
// and the row's L in Xhdr[row] (255: the row does not fit -- L > kXL -- and is read densely).  A DENSE row is the
// same thing with L = 32 and the identity code (slot `it` of chain j is element 16 * (it >> 1) + 2 * j + (it & 1) of
// the row), so the multiply loop is shared and only the loads differ.
// When a staged price column holds a non-finite value (the reference's own Q12 NaN cascade) 0 * p is NaN and the
// zeros matter: that replication's rows are all read densely that day, in the canonical order.
//
// Staged columns live in shared memory in a chain-major swizzle: the 16-byte unit holding columns 2*c2, 2*c2+1
// of stock m is unit ((e * C2 + c2) * 8 + j) with j = (m >> 1) & 7 (the stock's chain) and e = 2 * (m >> 4) +
// (m & 1) (its position in the chain).  Lane j of a row only ever gathers stocks of chain j, so the 8 lanes of a
// quarter warp always hit 8 different 16-byte bank groups: every gather is conflict-free, whatever the stocks.
//
// Who keeps the mirror: rows whose holdings changed since their packed copy was written are marked -- by
// events-1 in the day's modified-fund bitmap (hand_i: divested and respawned funds), by events-2 in the
// replication's `xdirty` bitmap (funds that received shares) -- and the pass re-packs the marked rows from the
// dense holdings before it multiplies (a walk over the row in chain order).  fsb_init_device / fsb_upload_state mark
// every row.
//
// Code size matters as much as bytes here (the step kernels are instruction-fetch bound; a first version
// of this kernel with the slots of a block unrolled in registers ran at a 65 % instruction-cache hit rate and was
// SLOWER than the dense pass): the multiply loops are rolled -- one chain slot per trip, ~(8 + 3 C2) instructions --
// and the packed rows reach the lanes through a per-warp ring of two shared-memory stages (a stage = the 4 rows of a
// trip, one bulk asynchronous copy of 128 + 64 L bytes per row completing on the stage's mbarrier), one trip ahead of
// the arithmetic (`prefetch.global.L1` + cached loads were tried first: 26 % L1 hit rate, every slot waited for DRAM);
// the chain sums are combined through a small per-warp shared-memory transpose (same additions as the butterfly:
// ((c0+c4)+(c2+c6)) + ((c1+c5)+(c3+c7))) instead of ~30 shuffles.
#pragma once
#include <cfloat>
#include "fsb_step.cuh"

namespace fsb {

constexpr int kXL = 16;                             // chain slots of a packed row
constexpr int kXIdxBytes = (kXL / 8) * 64;          // 128
constexpr int kXRowBytes = kXIdxBytes + kXL * 64;   // 1152
constexpr int kXThreads = 128;                      // 4 warps; several CTAs per SM so that one replication's set-up hides behind another's rows
constexpr int kXCols = 12;                          // staged columns per chunk (first chunk: 2 price columns + 10 horizons)
constexpr int kXKc = 128;                           // horizon list capacity (DevState::kc <= W <= ... checked by the host)

constexpr int kXPlane = 4 * 9 + 14;        // doubles per chain plane of the per-warp epilogue scratch: rows at pitch 9, planes 2 (mod 16) apart
constexpr int kXSlot = kXRowBytes + 64;      // bytes of a row slot of the ring (odd multiple of 64: the 4 slots of a stage alternate bank halves)
constexpr int kXStage = 4 * kXSlot;          // a stage = the 4 rows of a trip
constexpr int kXWarps = kXThreads >> 5;

#ifdef FSB_PHASE_TIMING
#define XMARK(id) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_phase_cycles[id], static_cast<unsigned long long>(now_ - sm.tmark)); sm.tmark = now_; } } while (0)
#else
#define XMARK(id) do { } while (0)
#endif

struct XShared {
    double cols[kXCols * 256];   // swizzled staged columns (24 KB)
    double nav[256];             // the day's NAV at the post-impact prices (numerator of every horizon return)
    unsigned long long bars[kXWarps * 2];  // mbarriers of the ring stages
    unsigned short order[272];   // row order of the chunk: by descending chain length (dense rows first), 0xffff: none
    unsigned short marked[256];  // rows to re-pack
    unsigned char hdr[256];
    int hcols[kXKc];
    unsigned fmod[8], dirty[8];
    int hist[36], cursor[36];
    int nrows, nmarked;
    long long r, rnext, tmark;
    __align__(128) unsigned char ring[kXWarps * 2 * kXStage];  // per warp: two stages of packed rows
};

__device__ __forceinline__ int imax8(int a)
{
    a = max(a, __shfl_xor_sync(0xffffffffu, a, 4));
    a = max(a, __shfl_xor_sync(0xffffffffu, a, 2));
    a = max(a, __shfl_xor_sync(0xffffffffu, a, 1));
    return a;
}

// re-pack the marked rows of the replication from the dense holdings: a quarter warp per row, lane j walks chain j
static __device__ __noinline__ void passx_repack(unsigned char *X, unsigned char *Xh, const double *H, XShared &sm)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 7, q = lane >> 3;
    const double2 *const H2 = reinterpret_cast<const double2 *>(H);
#pragma unroll 1
    for (int b = 4 * warp; b < sm.nmarked; b += 4 * kXWarps) {
        const bool on = b + q < sm.nmarked;
        const int k = on ? static_cast<int>(sm.marked[b + q]) : 0;
        unsigned char *rp = X + static_cast<long long>(k) * kXRowBytes + 2 * j;
        double *vp = reinterpret_cast<double *>(X + static_cast<long long>(k) * kXRowBytes + kXIdxBytes) + 2 * j;
        const double2 *hp = H2 + static_cast<long long>(k) * 128 + j;
        int pos = 0;  // slot `pos` of this lane's chain lives at offset 16 * (pos >> 1) + (pos & 1) from rp / vp
        double2 h[16];  // the whole chain first: the byte stores below alias everything, loads after them would be serialised
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = on ? __ldcg(hp + 8 * i) : make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (h[i].x != 0.0) {  // (a NaN holding counts as present)
                if (pos < kXL) { rp[16 * (pos >> 1) + (pos & 1)] = static_cast<unsigned char>(2 * i); vp[16 * (pos >> 1) + (pos & 1)] = h[i].x; }
                ++pos;
            }
            if (h[i].y != 0.0) {
                if (pos < kXL) { rp[16 * (pos >> 1) + (pos & 1)] = static_cast<unsigned char>(2 * i + 1); vp[16 * (pos >> 1) + (pos & 1)] = h[i].y; }
                ++pos;
            }
        }
        const int L = (imax8(pos) + 1) & ~1;  // whole pairs
        if (on) {
            for (int p = pos; p < min(L, kXL); ++p) { rp[16 * (p >> 1) + (p & 1)] = 0; vp[16 * (p >> 1) + (p & 1)] = 0.0; }  // pad: (first stock of the chain, 0.0)
            if (j == 0) {
                const unsigned char hv = (L > kXL) ? 255 : static_cast<unsigned char>(L);
                sm.hdr[k] = hv;
                Xh[k] = hv;
            }
        }
    }
}

// one chunk of staged columns against every row of the replication.
//   C2      column pairs staged (columns 2*C2)
//   hb      first staged slot that holds a horizon column (2 in the first chunk, 0 later)
//   nh      horizon columns of this chunk, kcol0: index of the first of them in the key rows
//   first   first chunk of the day: slot 0 = opening prices, slot 1 = post-impact prices
//   alldense  a staged price is not finite: every row is read densely (zeros included)
//   gcount  stages this warp's ring has carried so far (slot and phase parity of a stage follow from its running index)
template <int C2>
__device__ __forceinline__ void passx_rows(const DevState &s, XShared &sm, const int hb, const int nh, const int kcol0, const int want_keys,
                                            const bool first, const int open_day, const int alldense, unsigned &gcount)
{
    constexpr int NC = 2 * C2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 7, q = lane >> 3;
    const long long r = sm.r;
    const int K = s.K, kq = (K + 3) & ~3;
    const unsigned char *const X = s.X + (r * kq) * static_cast<long long>(kXRowBytes);
    const unsigned char *const Hb = reinterpret_cast<const unsigned char *>(s.H + r * s.h_stride);
    const double2 *const cols2 = reinterpret_cast<const double2 *>(sm.cols) + j;
    double *const nav_open = s.nav_open + r * K, *const nav_close = s.nav_close + r * K;
    unsigned long long *const keys = s.keys + r * (static_cast<long long>(s.kc) * kq);
    unsigned char *const ring = sm.ring + warp * (2 * kXStage);
    unsigned long long *const bars = sm.bars + warp * 2;
    const int ntrips = (sm.nrows + 3) >> 2;
    // Lanes 0..3 look after the four rows of a trip (lane l: quarter l): row id and chain length from the order list, the
    // row's bulk copy into stage g of the ring (completing on the stage's mbarrier), and they hand (row, length) to their
    // quarter when the trip is multiplied.  Every lane is done with the slot's previous contents (__syncwarp): their
    // shared-memory loads have returned, so the asynchronous proxy may overwrite it (write-after-read needs no proxy fence;
    // the generic writes of the epilogue scratch are fenced).
    int f_row = 0xffff, f_h = 0;  // (lanes 0..3) the row the last fill fetched for quarter `lane`
    auto fill = [&](int trip, unsigned g) {
        fence_proxy_async();
        __syncwarp();
        unsigned bytes = 0;
        if (lane < 4) {
            f_row = (4 * trip + lane < sm.nrows) ? static_cast<int>(sm.order[4 * trip + lane]) : 0xffff;
            f_h = (f_row != 0xffff) ? static_cast<int>(sm.hdr[f_row]) : 0;
            bytes = (f_row == 0xffff || alldense || f_h == 255 || f_h == 0) ? 0u : static_cast<unsigned>(kXIdxBytes + 64 * f_h);
        }
        unsigned total = bytes + __shfl_xor_sync(0xffffffffu, bytes, 1);
        total += __shfl_xor_sync(0xffffffffu, total, 2);
        if (lane == 0) mbar_expect_tx(bars + (g & 1u), total);
        __syncwarp();  // (expect_tx is in place before any copy can complete)
        if (bytes) bulk_g2s(ring + (g & 1u) * kXStage + lane * kXSlot, X + static_cast<long long>(f_row) * kXRowBytes, bytes, bars + (g & 1u));
    };
    const int my_trips = (ntrips - warp + kXWarps - 1) / kXWarps;
    if (my_trips <= 0) return;
    const unsigned g0 = gcount;
    fill(warp, g0);
#pragma unroll 1
    for (int i = 0; i < my_trips; ++i) {
        const int trip = warp + i * kXWarps;
        const unsigned g = g0 + static_cast<unsigned>(i);
        const int row = __shfl_sync(0xffffffffu, f_row, q), h = __shfl_sync(0xffffffffu, f_h, q);  // (before the next fill replaces them)
        if (i + 1 < my_trips) fill(trip + kXWarps, g + 1u);
        const bool on = row != 0xffff;
        const bool dense = on && (alldense || h == 255);
        double acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = 0.0;
        if (__any_sync(0xffffffffu, dense)) {
            // ---- dense rows (rare): slot `it` of chain j is element 16 (it >> 1) + 2 j + (it & 1) of the row, straight from HBM ----
            const double2 *hp = reinterpret_cast<const double2 *>(Hb + static_cast<long long>(on ? row : 0) * 2048) + j;
#pragma unroll 4
            for (int ii = 0; ii < 16; ++ii) {
                if (dense) {
                    const double2 hv = __ldcg(hp + 8 * ii);
                    const double2 *cp0 = cols2 + (2 * ii) * (C2 * 8), *cp1 = cp0 + C2 * 8;
#pragma unroll
                    for (int c2 = 0; c2 < C2; ++c2) {
                        const double2 p0 = cp0[c2 * 8], p1 = cp1[c2 * 8];
                        acc[2 * c2] = fma(hv.x, p0.x, acc[2 * c2]);
                        acc[2 * c2] = fma(hv.y, p1.x, acc[2 * c2]);
                        acc[2 * c2 + 1] = fma(hv.x, p0.y, acc[2 * c2 + 1]);
                        acc[2 * c2 + 1] = fma(hv.y, p1.y, acc[2 * c2 + 1]);
                    }
                }
            }
        }
        XMARK(29);
        mbar_wait(bars + (g & 1u), (g >> 1) & 1u);
        XMARK(30);
        {
            // ---- packed rows: lane j walks chain j of the quarter's row in the stage, two slots per step ------------
            const int L = (on && !dense) ? h : 0;
            int Lmax = max(L, __shfl_xor_sync(0xffffffffu, L, 8));
            Lmax = max(Lmax, __shfl_xor_sync(0xffffffffu, Lmax, 16));
            const unsigned char *slot = ring + (g & 1u) * kXStage + q * kXSlot;
            const unsigned short *is = reinterpret_cast<const unsigned short *>(slot) + j;
            const double2 *vs = reinterpret_cast<const double2 *>(slot + kXIdxBytes) + j;
#pragma unroll 1
            for (int it = 0; it < Lmax; it += 2) {
                if (it < L) {
                    const double2 v = vs[it * 4];
                    const unsigned ee = is[it * 4];
                    const double2 *cp0 = cols2 + (ee & 0xffu) * (C2 * 8), *cp1 = cols2 + (ee >> 8) * (C2 * 8);
#pragma unroll
                    for (int c2 = 0; c2 < C2; ++c2) {
                        const double2 p0 = cp0[c2 * 8], p1 = cp1[c2 * 8];
                        acc[2 * c2] = fma(v.x, p0.x, acc[2 * c2]);
                        acc[2 * c2 + 1] = fma(v.x, p0.y, acc[2 * c2 + 1]);
                        acc[2 * c2] = fma(v.y, p1.x, acc[2 * c2]);
                        acc[2 * c2 + 1] = fma(v.y, p1.y, acc[2 * c2 + 1]);
                    }
                }
            }
        }
        // ---- epilogue: the 8 chain sums of every (row, column) -> one sum, through a scratch [chain][row of the trip * 9 +
        //      column of the group], 8 columns at a time; lane (qq, cc) then owns column cg + cc of the trip's row qq.
        //      The scratch is the stage just consumed (free until the fill at the top of the next trip) ---------------
        double *const red = reinterpret_cast<double *>(ring + (g & 1u) * kXStage);
        static_assert(8 * kXPlane * 8 <= kXStage, "the epilogue scratch must fit in a ring stage");
        XMARK(29);
        const int qq = q, cc = j, orow = row;  // (lane (q, j) owns column cg + j of its own quarter's row)
        const int k = on ? row : 0;
#pragma unroll
        for (int cg = 0; cg < NC; cg += 8) {
            const int G = (NC - cg >= 8) ? 8 : NC - cg;  // columns of this group (8 or 4)
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < G) red[j * kXPlane + q * 9 + c] = acc[(cg + c) % NC];
            __syncwarp();
            double sum = 0.0;
            if (cc < G) {
                const double *rd = red + qq * 9 + cc;
                const double c0 = rd[0], c1 = rd[kXPlane], c2 = rd[2 * kXPlane], c3 = rd[3 * kXPlane], c4 = rd[4 * kXPlane],
                             c5 = rd[5 * kXPlane], c6 = rd[6 * kXPlane], c7 = rd[7 * kXPlane];
                sum = ((c0 + c4) + (c2 + c6)) + ((c1 + c5) + (c3 + c7));
            }
            const int col = cg + cc;
            const bool mine = (cc < G) && (orow != 0xffff);
            if (first && cg == 0) {
                if (mine) {
                    if (col == 0) {
                        if (!((sm.fmod[k >> 5] >> (k & 31)) & 1u)) nav_open[k] = sum;
                        if (open_day) nav_close[k] = sum;
                    } else if (col == 1 && !open_day) {
                        sm.nav[k] = sum;
                        nav_close[k] = sum;
                    }
                }
                __syncwarp();
            }
            const int hc = col - hb;
            if (mine && hc >= 0 && hc < nh) {
                const double navk = sm.nav[k];
                const double rho = (navk - sum) / sum;
                const unsigned long long word = want_keys ? sort_key(rho) : static_cast<unsigned long long>(__double_as_longlong(rho));
                __stcg(keys + static_cast<long long>(kcol0 + hc) * kq + k, word);
            }
        }
        XMARK(31);
    }
    gcount = g0 + static_cast<unsigned>(my_trips);
}

__global__ void __launch_bounds__(kXThreads, 3) fsb_passx_kernel(const DevState s, const StepArgs a, const int launch_parity)
{
    XShared &sm = *reinterpret_cast<XShared *>(fsb_smem);  // dynamic shared memory (sizeof(XShared) > the 48 KB static limit)
    const int tid = threadIdx.x;
    const int K = s.K, nf = (K + 31) / 32, t = a.t_first;
    int *const queue = s.queue + 2 * (4 * a.batch + KIND_PASS);
    if (blockIdx.x == 0 && tid == 0) queue[launch_parity ^ 1] = 0;  // reset the next launch's counter
    if (tid == 0) {
        sm.rnext = a.r_begin + atomicAdd(&queue[launch_parity], 1);
        for (int b = 0; b < kXWarps * 2; ++b) mbar_init(sm.bars + b, 1);  // the ring's mbarriers, once per kernel
        fence_mbar_init();
    }
    unsigned gcount = 0;  // stages this warp's ring has carried (uniform in the warp)
#ifdef FSB_PHASE_TIMING
    if (tid == 0) sm.tmark = clock64();
#endif
    for (;;) {
        __syncthreads();  // (the previous replication's readers of sm are done; rnext is visible)
        const long long r = sm.rnext;
        if (r >= a.r_end) break;
        __syncthreads();
        if (tid == 0) {  // the next replication is fetched while this one is worked on (the atomic's latency is off the critical path)
            sm.r = r;
            sm.rnext = a.r_begin + atomicAdd(&queue[launch_parity], 1);
        }
        if (s.err[r] & ~FSB_REP_TRACE_OVERFLOW) continue;  // frozen replication (uniform)
        const int *hi = s.hand_i + r * s.hi_stride;
        if (hi[HI_T] != t) continue;                        // (events-1 stopped this replication: its error flag is set)
        const int open_day = hi[HI_OPEN], ncols = open_day ? 0 : hi[HI_NCOLS];
        if (tid < 8) {
            sm.fmod[tid] = (tid < nf) ? static_cast<unsigned>(hi[HI_FMOD + tid]) : 0u;
            sm.dirty[tid] = (tid < nf) ? s.xdirty[r * 8 + tid] : 0u;
        }
        for (int c = tid; c < ncols; c += kXThreads) sm.hcols[c] = hi[HI_FMOD + nf + c];
        for (int k = tid; k < 256; k += kXThreads) sm.hdr[k] = s.Xhdr[r * 256 + k];
        const double *P = s.P + r * s.p_stride;
        const int want_keys = (a.selector == FSB_SEL_PROBABILISTIC) ? 1 : 0;
        if (tid == 0) sm.nmarked = 0;
        __syncthreads();
        XMARK(26);
        // ---- re-pack the rows whose holdings changed since their packed copy was written ---------------------------
        for (int k = tid; k < K; k += kXThreads)
            if (((sm.dirty[k >> 5] | sm.fmod[k >> 5]) >> (k & 31)) & 1u) sm.marked[atomicAdd(&sm.nmarked, 1)] = static_cast<unsigned short>(k);
        __syncthreads();
        if (sm.nmarked > 0) {
            passx_repack(s.X + (r * ((K + 3) & ~3)) * static_cast<long long>(kXRowBytes), s.Xhdr + r * 256, s.H + r * s.h_stride, sm);
            __syncthreads();  // (the packed rows written above are read below with L2 loads by other warps of this CTA)
        }
        XMARK(27);
        // chunks of horizon columns: the first holds the two price columns + up to xcols - 2 horizons (xcols = kXCols unless FSB_XCOLS)
#pragma unroll 1
        for (int c0 = 0; c0 == 0 || c0 < ncols; c0 += (c0 == 0 ? s.xcols - 2 : s.xcols)) {
            const bool first = (c0 == 0);
            const int hb = first ? 2 : 0;
            const int nh = min(s.xcols - hb, ncols - c0);
            const int nc = hb + nh;                                   // staged columns
            const int C2 = (nc <= 4) ? 2 : (nc <= 8) ? 4 : 6;
            // ---- stage the columns (a thread owns two stocks), chain-major swizzle ------------------------------
            bool bad = false;
#pragma unroll 2
            for (int c = 0; c < 2 * C2; ++c) {
                const double *src;
                if (first && c == 0) src = s.hand_p + r * s.Mp;
                else if (first && c == 1) src = open_day ? s.hand_p + r * s.Mp : P + static_cast<long long>(pcol(s, t)) * s.Mp;
                else if (c < nc) src = P + static_cast<long long>(pcol(s, t - sm.hcols[c0 + c - hb])) * s.Mp;
                else src = s.hand_p + r * s.Mp;  // spare slot: finite data, results dropped
                const double2 v = __ldcg(reinterpret_cast<const double2 *>(src) + tid);  // stocks 2 tid, 2 tid + 1: chain tid & 7, e = 2 (tid >> 3) + {0, 1}
                bad = bad || !(fabs(v.x) <= DBL_MAX) || !(fabs(v.y) <= DBL_MAX);
                const int jj = tid & 7, e0 = (tid >> 3) << 1;
                sm.cols[((((e0) * C2 + (c >> 1)) * 8 + jj) << 1) + (c & 1)] = v.x;
                sm.cols[((((e0 + 1) * C2 + (c >> 1)) * 8 + jj) << 1) + (c & 1)] = v.y;
            }
            if (tid < 36) sm.hist[tid] = 0;
            for (int o = tid; o < 272; o += kXThreads) sm.order[o] = 0xffff;
            const int nonfinite = __syncthreads_or(bad ? 1 : 0);
            // ---- row order: by descending chain length, dense rows (32 slots) first (counting sort) ----------------
            int bucket[2] = {-1, -1};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = tid + h * kXThreads;
                if (k < K) {
                    const int L = (nonfinite || sm.hdr[k] == 255) ? 32 : min(static_cast<int>(sm.hdr[k]), kXL);
                    bucket[h] = 32 - L;
                    atomicAdd(&sm.hist[bucket[h]], 1);
                }
            }
            __syncthreads();
            if (tid == 0) {
                int off = 0;
                for (int b = 0; b < 33; ++b) { sm.cursor[b] = off; off += sm.hist[b]; }
                sm.nrows = off;
            }
            __syncthreads();
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (bucket[h] >= 0) sm.order[atomicAdd(&sm.cursor[bucket[h]], 1)] = static_cast<unsigned short>(tid + h * kXThreads);
            __syncthreads();
            XMARK(28);
            switch (C2) {
            case 2: passx_rows<2>(s, sm, hb, nh, c0, want_keys, first, open_day, nonfinite, gcount); break;
            case 4: passx_rows<4>(s, sm, hb, nh, c0, want_keys, first, open_day, nonfinite, gcount); break;
            default: passx_rows<6>(s, sm, hb, nh, c0, want_keys, first, open_day, nonfinite, gcount); break;
            }
            __syncthreads();  // (the next chunk overwrites the staged columns and the row order)
            XMARK(29);
        }
        if (tid < nf) s.xdirty[r * 8 + tid] = 0u;
        const int slot = s.trace_slot ? s.trace_slot[r] : -1;
        if (slot >= 0) {  // trace point of a traced replication: the NAVs of :460 (cold)
            double *tr = s.tr_nav + static_cast<long long>(slot) * K * s.T + static_cast<long long>(K) * (t - 1);
            for (int k = tid; k < K; k += kXThreads) tr[k] = __ldcg(s.nav_open + r * K + k);
        }
    }
}

}  // namespace fsb
