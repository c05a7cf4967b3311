// fsb_state.h -- structure-of-arrays device state of the engine (replication x fund x asset).
// One DevState is passed by value to every kernel.  See DESIGN.md "Data layout in HBM".
#pragma once
#include <cstdint>
#include "../../include/fsb.h"

namespace fsb {

struct DevPoint {  // one sweep point = one Params (src/params.jl:4-31), shapes excluded
    double marketstartval, drift, marketvol, betamean, betastd, stockstartval;
    double reviewprob, thresholdmean, thresholdstd;
    int perf_lo, perf_hi, cap_lo, cap_hi, psize_lo, psize_hi;
    int n_impact, n_vol, impact_off, vol_off;  // offsets into DevState::tables
};

constexpr int kFlowBins = FSB_STATS_FLOW_BINS;
constexpr int kCnt = 8;  // per-replication counters, see CNT_*
enum { CNT_REVIEWS = 0, CNT_DIVEST = 1, CNT_REINVEST = 2, CNT_STEPS = 3, CNT_FLOOR = 4, CNT_RESPAWN = 5 };

struct DevState {
    // shapes
    int K, M, N, T, W;
    int Mp;         // row pitch of holdings / price columns: M rounded up to a multiple of 128 (zero padded)
    int hist_cols;  // price columns kept per replication: T (keep_history) or 2(W+1)
    int mkt_len;    // T (keep_history) or 2
    int keep_history;
    int cap_ev, cap_log, cap_trace, cmax;
    int kc;         // key columns (distinct reinvestor horizons) a replication can have per step
    int hi_stride;  // ints per replication in hand_i
    int slots_a, slots_c;  // shared-memory column slots of the events-1 / events-2 kernels
    int big;               // shapes whose per-asset arrays do not fit in shared memory: the FSB_BIG build of the step kernels
    long long big_stride;  // bytes of a CTA's global buffer for those arrays
    unsigned char *bigbuf; // [2 * batches][scratch_ctas][big_stride]
    int scan_choice;       // FSB_SCAN_CHOICE=1: the rank rule is sampled by the floating-point scan only (no integer shortcut; tests)
    int pass16;            // default shape: chunks of <= 8 columns take the 4 x 1 arrangement (full_pass_tma16); FSB_PASS16=0: never
    int pass_ahead;        // default shape: warp 0 of a pass CTA prepares the CTA's next replication during the current one (PassAhead); FSB_PASS_AHEAD=0: off
    int smem_sort;         // FSB_SMEM_SORT=1: 256 < K <= 2048 ranks by the shared-memory network only (no register network; tests)
    int ring_stages;       // per-warp holdings ring of the pass kernel (kRingStages on the default shape, else 0)
    int pass_parts;        // big shapes, tiled pass: a replication's rows are split into this many work units (>= 1)
    int pass_chunks;       // ... and a day's column chunks into this many (>= 1; 1: every work unit loops over all chunks)
    int n_points;
    long long R, rep_offset, reps_per_point;
    long long h_stride;  // doubles per replication in H
    long long p_stride;  // doubles per replication in P
    unsigned long long seed;
    // fund x asset
    double *H;  // [R][K4][Mp] holdings, asset fastest; K4 = K rounded up to 4 (zero rows, never written)
    // packed mirror of the holdings rows, read by fsb_passx_kernel (fsb_passx.cuh); null: the dense pass is used
    unsigned char *X;     // [R][K4][kXRowBytes]
    unsigned char *Xhdr;  // [R][256] chain length L of every packed row (255: row is read densely)
    int xcols;            // staged columns per chunk of that kernel (kXCols; FSB_XCOLS=4..16 for tests)
    unsigned *xdirty;     // [R][8] funds whose row changed after the day's pass (events-2); the next pass re-packs them
    // asset
    double *P;       // [R][hist_cols][Mp] price history (burn-in 1..W+1, then ring or full)
    double *beta, *vol, *impact;  // [R][Mp]
    double *volume;  // [R][T][Mp] or null
    double *market;  // [R][mkt_len]
    // investor SoA (one position per investor: fund id, K = cash)
    int *pos, *horizon;                 // [R][N]
    double *stake, *amount, *threshold; // [R][N]
    // fund
    double *nav_open, *nav_close, *cap0;  // [R][K]
    unsigned char *hzero;                 // [R][K] holdings row is all zero
    unsigned char *stakeless;             // [R][K] sum(stakes[k,:]) == 0 (respawn trigger, functions.jl:471)
    int *live_row;                        // [R][K] log row currently describing slot k
    // liquidation log, functions.jl:452-456
    int *log_n;                                              // [R]
    int *log_liquidated, *log_replacedby, *log_failtime, *log_ninv, *log_nstocks;  // [R][cap_log]
    double *log_size, *log_liquidity;                        // [R][cap_log]
    // status / statistics inputs
    int *err;         // [R] FSB_REP_* bits
    int *cnt;         // [R][kCnt]
    double *mkt_peak, *mkt_dd;  // [R]
    int *flow_hist;   // [R][kFlowBins]
    // tapes (null = Philox)
    const double *const *z_mkt, *const *z_stock, *const *u_review;  // [R] device pointers
    const fsb_tape_rec *const *ev_tape;                              // [R]
    const long long *ev_n;                                           // [R]
    int any_tape;
    // batched one-step tape fed from host buffers each step (fsb_run_step_hosttape): [R], [R][M], [R][N]
    const double *st_zmkt, *st_zstock, *st_urev;
    int step_tape;
    // trace
    const int *trace_slot;  // [R] -1 = not traced
    double *tr_nav, *tr_psell, *tr_presp;  // [slots][K*T], [slots][M*T]
    fsb_trace_rec *tr_recs;                // [slots][cap_trace]
    int *tr_n;                             // [slots]
    // per-CTA scratch + work queue
    double *scratch;  // [batches * 3][scratch_ctas][cap_ev][Mp] order rows beyond the shared-memory ones
    int scratch_ctas;
    unsigned long long *keys;  // [R][kc][K4] horizon returns of the step being processed (pass kernel -> events-2 kernel)
    // handoff between the three kernels of a step (fsb_step.cuh)
    int *hand_i;               // [R][hi_stride]
    double *hand_p;            // [R][Mp] opening prices of the day
    unsigned *hand_m;          // [R][2 * ceil(N / 64)] reviewer flags of the day (fsb_draws_kernel -> events-1), or null
    double *hand_s;            // [R][Mp] net sale per asset of the day (keep_history only, else null)
    const double *ranktab;     // [K] (i + 1) / (K (K + 1) / 2): rank probabilities of the rank rule (Q14)
    // per-replication pointers of the step context: [kRepPtrs] x (base address, stride in bytes); lane i of a
    // CTA computes pointer i of the replication it starts on (order = the pointer block of Ctx, fsb_step.cuh)
    const ulonglong2 *ptr_table;
    int *queue;       // [batches][4][2] dynamic work counters per kernel kind (double buffered across launches)
    const DevPoint *points;
    const double *tables;
};

__host__ __device__ inline int pcol(const DevState &s, int t)  // 1-based t -> column slot
{
    if (s.keep_history || t <= s.W + 1) return t - 1;
    return (s.W + 1) + ((t - 1) % (s.W + 1));
}
__host__ __device__ inline int mslot(const DevState &s, int t) { return s.keep_history ? t - 1 : (t & 1); }

}  // namespace fsb
