/*
 * fsb.h -- C ABI of libfsb_b200.so, the B200-native engine for the FundStabABM.jl simulation step.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own; the seam is
 * the function pair
 *     Func.initialise(market, stocks, investors, funds, params)            src/functions.jl:417
 *     Func.modelrun(market, stocks, investors, funds, params, fundselector) src/functions.jl:446
 * called from runmodel (src/runmodel.jl:30,36).  A Julia maintainer binds these entry points with
 * `ccall` (see INTEGRATION.md and julia/FundStabB200.jl); the Python mirror in
 * fundstababm.jl_b200/ binds the same symbols with ctypes.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.  Every function returns an
 *     int32 status (FSB_OK == 0, negative = error) unless stated otherwise; fsb_last_error()
 *     returns a thread-local, library-owned message valid until the next call on that thread.
 *   - All host arrays are COLUMN-MAJOR exactly as Julia stores the structs of src/types.jl:15-50
 *     (A[i,j] at (i-1) + rows*(j-1)); investor/fund/stock ids inside arrays and logs are 1-based
 *     as in Julia; time t is 1-based.  The caller owns every host array; the library copies in/out
 *     during the call and never retains host pointers.  The library owns all device memory.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef FSB_H
#define FSB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSB_VERSION 200 /* 0.2.0: fsb_config lost steps_per_launch; days must be run in order */

/* ---- status codes ---------------------------------------------------------------------- */
enum {
    FSB_OK = 0,
    FSB_ERR_BAD_PARAMS = -1,   /* inconsistent Params / config / arguments                    */
    FSB_ERR_BAD_SELECTOR = -2, /* Q9: selector not one of best/goodenough/probabilistic/random */
    FSB_ERR_CUDA = -3,         /* CUDA runtime failure (message has the CUDA error string)     */
    FSB_ERR_NCCL = -4,         /* NCCL failure or libnccl not loadable                         */
    FSB_ERR_BAD_STATE = -5,    /* uploaded state violates the one-position-per-investor form   */
    FSB_ERR_REPLICATION = -6,  /* at least one replication raised an error flag (see below)    */
    FSB_ERR_NO_HISTORY = -7,   /* operation needs keep_history=1                               */
    FSB_ERR_RANGE = -8         /* replication index / time range out of bounds                 */
};

/* per-replication error flags (bit set), fsb_rep_status() */
enum {
    FSB_REP_OK = 0,
    FSB_REP_Q6_NO_ANCHOR = 1,     /* more dead funds than cash holders (functions.jl:282 throws) */
    FSB_REP_TAPE_MISSING = 2,     /* tape mode: a needed draw is not on the tape               */
    FSB_REP_TAPE_INVALID = 4,     /* tape mode: anchor on the tape is not a cash holder         */
    FSB_REP_EVENT_OVERFLOW = 8,   /* more reviewers/orders in a step than event_capacity       */
    FSB_REP_LOG_OVERFLOW = 16,    /* liquidation log needs more rows than log_capacity         */
    FSB_REP_TRACE_OVERFLOW = 32   /* trace buffer full (trace only; simulation continues)      */
};

/* fundselector, src/functions.jl:344-353 */
enum { FSB_SEL_BEST = 0, FSB_SEL_GOODENOUGH = 1, FSB_SEL_PROBABILISTIC = 2, FSB_SEL_RANDOM = 3 };

/* fsb_run flags */
enum {
    FSB_RUN_NO_VOLUME_QUIRK = 1 /* record sell volume once per step instead of Q3's 1-3 times */
};

/* ---- Params.default, src/params.jl:4-31, as a POD (ranges flattened; plotpath is host-only) */
typedef struct fsb_params {
    int64_t bigm, bign, bigt, bigk;
    double marketstartval, drift, marketvol;
    const double *impact_values; /* collect(impactrange) */
    int64_t n_impact;
    double betamean, betastd, stockstartval;
    const double *vol_values;    /* collect(stockvolrange) */
    int64_t n_vol;
    double reviewprobability;
    int64_t perf_lo, perf_hi;    /* perfwindow   */
    int64_t cap_lo, cap_hi;      /* invcaprange  */
    double thresholdmean, thresholdstd;
    int64_t psize_lo, psize_hi;  /* portfsizerange */
} fsb_params;

/* ---- engine configuration -------------------------------------------------------------- */
typedef struct fsb_config {
    int64_t n_reps;         /* replications held by this engine (its shard)                    */
    int64_t rep_offset;     /* global id of local replication 0 (Philox counters use global ids,
                               so results do not depend on how replications are sharded)        */
    int64_t reps_per_point; /* sweep: global replication g belongs to point g / reps_per_point;
                               0 = all replications use point 0                                 */
    int32_t device;         /* CUDA device ordinal                                              */
    int32_t keep_history;   /* 1: full M x T price, volume and market history per replication
                               (single-replication drop-in mode, needed by fsb_download_state);
                               0: burn-in columns 1..W+1 plus a ring of the last W+1 columns    */
    int32_t event_capacity; /* max reviewers / orders per replication-step; 0 = auto            */
    int32_t log_capacity;   /* liquidation-log rows per replication; 0 = auto (K + 1024)        */
    int32_t trace_capacity; /* trace records per traced replication; 0 = auto                   */
    int32_t block_threads;  /* 0 = auto                                                         */
    int32_t blocks_per_sm;  /* CTAs per SM of the holdings-pass kernel; 0 = auto                */
} fsb_config;

typedef struct fsb_engine fsb_engine;

/* trace / tape records (same numbering as the oracle, oracle/fsb_oracle.h) */
typedef struct fsb_trace_rec { int32_t t, kind, a, b; double v; } fsb_trace_rec;
typedef struct fsb_tape_rec { uint64_t key; double value; } fsb_tape_rec;
enum { FSB_EV_REVIEWER = 1, FSB_EV_DIVEST = 2, FSB_EV_CASHOUT = 3, FSB_EV_FLOOR = 4,
       FSB_EV_SPAWN = 5, FSB_EV_SPAWN_N = 6, FSB_EV_CHOICE = 7, FSB_EV_PICK = 8 };
enum { FSB_TAPE_ANCHOR = 1, FSB_TAPE_RPSIZE = 2, FSB_TAPE_RPPICK = 3, FSB_TAPE_CHOICE_U = 4,
       FSB_TAPE_CHOICE_IDX = 5 };

/* ---- ensemble statistics (integer-only, so the NCCL sum is order independent) ------------ */
#define FSB_STATS_HEADER 16
#define FSB_STATS_LIQ_BINS 256
#define FSB_STATS_DD_BINS 64
#define FSB_STATS_FLOW_BINS 64
enum {
    FSB_ST_REPS = 0, FSB_ST_LIQUIDATIONS, FSB_ST_FUNDS_EVER, FSB_ST_FLOOR_HITS, FSB_ST_DIVESTMENTS,
    FSB_ST_REVIEWS, FSB_ST_REINVESTMENTS, FSB_ST_ERR_REPS, FSB_ST_STEPS,
    FSB_ST_NAN_REPS /* replications whose prices went NaN (Q12: reinvestment into a zero-value fund) */
};
/* words per sweep point: header | failtime histogram [T+1] | liquidations-per-replication
 * histogram [256] | market max-drawdown histogram [64] | per-step redemption-flow histogram [64] */
int64_t fsb_stats_words_per_point(const fsb_engine *e);

/* ---- lifecycle --------------------------------------------------------------------------- */
int32_t fsb_version(void);
const char *fsb_last_error(void);
/* points[n_points]: one Params per sweep point; shapes (bigm,bign,bigt,bigk,perf_hi) must agree */
int32_t fsb_create(const fsb_params *points, int32_t n_points, const fsb_config *cfg, fsb_engine **out);
int32_t fsb_destroy(fsb_engine *e);

/* ---- state in: either upload the reference's structs (replaces nothing of Func.initialise:
 *      the caller ran it) or initialise every replication on the device (replaces
 *      Func.initialise, src/functions.jl:417-430, with Philox draws) -------------------------- */
int32_t fsb_upload_state(fsb_engine *e, int64_t rep,
                         const double *market /*T*/, const double *stock_value /*MxT*/,
                         const double *beta /*M*/, const double *vol /*M*/, const double *impact /*M*/,
                         const double *volume /*MxT*/, const double *inv_assets /*Nx(K+1)*/,
                         const int64_t *horizon /*N*/, const double *threshold /*N*/,
                         const double *holdings /*KxM*/, const double *stakes /*KxN*/,
                         const double *fund_value /*KxT, may be NULL*/);
int32_t fsb_init_device(fsb_engine *e, uint64_t seed);
/* Philox key for the step draws of an UPLOADED state (fsb_init_device sets it itself) */
int32_t fsb_set_seed(fsb_engine *e, uint64_t seed);

/* ---- equivalence mode: feed recorded draws instead of Philox (SURVEY.md 8c "tape") --------
 * dense tapes (any may be NULL): z_mkt[T], z_stock[MxT], u_review[NxT], column-major, 1-based t;
 * event tape: records sorted by key = fsb_tape_key(site,t,a,b), ids 0-based                     */
int32_t fsb_set_dense_tape(fsb_engine *e, int64_t rep, const double *z_mkt, const double *z_stock,
                           const double *u_review);
int32_t fsb_set_event_tape(fsb_engine *e, int64_t rep, const fsb_tape_rec *recs, int64_t n);
uint64_t fsb_tape_key(int32_t site, int64_t t, int64_t a, int64_t b);
/* record per-step NAVs, post-impact prices and the discrete event list of one replication */
int32_t fsb_set_trace(fsb_engine *e, int64_t rep, int32_t on);

/* ---- the hot path: replaces the time loop of Func.modelrun, src/functions.jl:457-489 --------
 * runs t = t_first..t_last (inclusive, 1-based; first valid t is perf_hi+2) on every
 * replication; blocking (returns after the stream is synchronised).  Builds the liquidation log
 * (functions.jl:448-456) on the first call.  Days are simulated once and in order: t_first must be
 * the day after the last simulated one (perf_hi+2 after fsb_init_device / fsb_upload_state of an
 * initialised state), else FSB_ERR_RANGE.  Returns FSB_ERR_REPLICATION when a replication raised
 * an error flag (the run itself completed; the flagged replications are frozen). */
int32_t fsb_run(fsb_engine *e, int64_t t_first, int64_t t_last, int32_t selector, int32_t flags);

/* batched equivalence mode, one step per call: this step's shocks for ALL replications come from
 * host buffers (z_mkt[R], z_stock[R][M] stock-fastest, u_review[R][N] investor-fastest; pinned
 * memory makes the copies asynchronous), are copied to the device, the step runs, and the NAVs
 * V[:,t] of every fund (:460) are copied back into nav_open_out[R][K] (may be NULL).  Blocking.
 * Like fsb_run_async + fsb_sync it does NOT scan the replication error flags (that would add a
 * device-to-host copy per step): callers check fsb_rep_status when the sequence of steps is done. */
int32_t fsb_run_step_hosttape(fsb_engine *e, int64_t t, int32_t selector, int32_t flags,
                              const double *z_mkt, const double *z_stock, const double *u_review,
                              double *nav_open_out);

/* ---- state out ----------------------------------------------------------------------------- */
int32_t fsb_download_state(fsb_engine *e, int64_t rep,
                           double *market, double *stock_value, double *beta, double *vol, double *impact,
                           double *volume, double *inv_assets, int64_t *horizon, double *threshold,
                           double *holdings, double *stakes, double *fund_value);
/* compact per-replication state that exists for any keep_history: holdings KxM (column-major),
 * current price column M, investor SoA (fund id 1-based, K+1 = cash), NAVs K; any may be NULL */
int32_t fsb_download_compact(fsb_engine *e, int64_t rep, double *holdings, double *price_now,
                             int64_t *inv_fund, double *inv_amount, double *inv_stake,
                             double *nav_open, double *nav_close, double *market_now);
/* liquidation log, columns of src/functions.jl:452-456; call with NULL columns to get n_rows */
int32_t fsb_download_log(fsb_engine *e, int64_t rep, int64_t *n_rows, int64_t *liquidated,
                         int64_t *replacedby, double *initialsize, double *liquidity,
                         int64_t *failtime, int64_t *ninvestors, int64_t *nstocks);
/* trace of a traced replication: nav_open KxT, p_sell MxT, p_resp MxT (column-major, zero where
 * not produced) and the event records; call with recs=NULL to get n_recs */
int32_t fsb_download_trace(fsb_engine *e, int64_t rep, double *nav_open, double *p_sell, double *p_resp,
                           fsb_trace_rec *recs, int64_t *n_recs);
int32_t fsb_rep_status(fsb_engine *e, int32_t *flags /*n_reps*/);

/* ---- stylised facts on the device ----------------------------------------------- */
/* Func.calc_stylisedfacts (src/functions.jl:604-648; called at src/runmodel.jl:47) for replications
 * [rep_first, rep_first + n_reps), computed from the price / volume / market histories in HBM
 * (keep_history=1), so that only M-vectors cross the host link instead of the M x T histories.
 * Array outputs of the reference function, per replication, column-major as Julia stores them
 * (any may be NULL): stock_pacf, stock_volacluster [n_reps][maxlag][M] (calc_pacf :588-602 of the
 * demeaned returns :517-535 / of their squares); market_pacf, market_volacluster [n_reps][maxlag];
 * kurtoses (:538-545), lossgain (:572-586, percentile `returnspctile` of |x|), corrs (:560-570)
 * [n_reps][M].  maxlag in 1..10 (reference default 10), returnspctile in [0,100] (default 5).
 * The z-tests and the KS test of :629-641 are host statistics on these vectors / on downloaded
 * returns and stay with the caller. */
int32_t fsb_stylised_facts(fsb_engine *e, int64_t rep_first, int64_t n_reps, int32_t maxlag, double returnspctile,
                           double *stock_pacf, double *market_pacf, double *stock_volacluster,
                           double *market_volacluster, double *kurtoses, double *lossgain, double *corrs);

/* ---- ensemble statistics + the single NCCL allreduce ---------------------------------------- */
/* multi-process: rank 0 calls fsb_comm_unique_id, ships the 128 bytes to the other ranks by any
 * means, then every rank calls fsb_comm_init_rank.  single-process multi-GPU: fsb_comm_init_all. */
int32_t fsb_comm_unique_id(uint8_t id[128]);
int32_t fsb_comm_init_rank(fsb_engine *e, int32_t n_ranks, int32_t rank, const uint8_t id[128]);
int32_t fsb_comm_init_all(fsb_engine **engines, int32_t n);
/* computes this engine's statistics on the device, sums them across the communicator (if any;
 * ncclAllReduce(sum, uint64) on the engine's stream) and copies n_points*words_per_point words */
int32_t fsb_stats(fsb_engine *e, uint64_t *out);
/* the same words for THIS engine's replications only: no collective (what fsb_stats sums over the ranks) */
int32_t fsb_stats_local(fsb_engine *e, uint64_t *out);
int32_t fsb_stats_all(fsb_engine **engines, int32_t n, uint64_t *out);

/* ---- measurement hooks (bench.py) ----------------------------------------------------------- */
/* device time of the last fsb_run (CUDA events): total on the engine's stream, and the time summed over
 * the launches of the dominant kernel (the holdings pass, fsb_pass_kernel; events on the stream of every
 * launch); number of kernel launches of the last fsb_run (five per simulated day and replication batch; three for
 * shapes beyond shared memory, where the draws and the fund choice stay inside the event kernels) */
int32_t fsb_last_run_timing(fsb_engine *e, double *total_ms, double *step_kernel_ms, int64_t *launches);
/* how many launches of the dominant kernel step_kernel_ms sums over */
int64_t fsb_last_pass_launches(const fsb_engine *e);
/* device time (CUDA events on the engine's stream) of the kernel launches of the last fsb_stylised_facts */
int32_t fsb_last_facts_timing(fsb_engine *e, double *kernel_ms);
/* asynchronous variant used for overlapping several engines in one process.  Ordering contract: the
 * engine's streams are non-blocking; every upload / download / tape / status entry point first waits for
 * the work enqueued here, so it is safe to call them after fsb_run_async without fsb_sync.  fsb_sync
 * does not scan the replication error flags (fsb_rep_status does). */
int32_t fsb_run_async(fsb_engine *e, int64_t t_first, int64_t t_last, int32_t selector, int32_t flags);
int32_t fsb_sync(fsb_engine *e);
int64_t fsb_device_bytes(const fsb_engine *e);
/* engines created under FSB_REDZONE=1 keep guard bands around every device buffer: counts the buffers some kernel
 * wrote outside of (n_buffers = -1 without red zones) */
int32_t fsb_debug_redzones(fsb_engine *e, int64_t *n_buffers, int64_t *n_damaged);
/* (tests) the column chunks the big-shape pass kernel makes of a day with ncols distinct reinvestor horizons
 * (src/functions.jl:335-379 asks for one horizon return per fund and distinct horizon): first[i] / count[i] horizons of
 * chunk i; returns the number of chunks, -1 if it exceeds cap */
int32_t fsb_debug_pass_chunks(int32_t ncols, int32_t *first, int32_t *count, int32_t cap);
/* the day fsb_run / fsb_run_async / fsb_run_step_hosttape must start with next (perf_hi+2 on a fresh state) */
int64_t fsb_next_day(const fsb_engine *e);
/* RNG building blocks evaluated ON THE DEVICE, for cross-checks against the oracle */
int32_t fsb_debug_draws(fsb_engine *e, uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b,
                        uint32_t idx0, int32_t n, int32_t kind /*0 normal 1 uniform 2 discrete*/,
                        uint32_t n_discrete, double *out);
int32_t fsb_debug_dot(fsb_engine *e, const double *x, const double *y, int64_t n, double *out);
/* cycles thread 0 of every CTA spent in each phase of the step kernel (profiling builds with
 * -DFSB_PHASE_TIMING only; zeros otherwise); reset != 0 clears the counters */
int32_t fsb_debug_phase_cycles(fsb_engine *e, uint64_t out[32], int32_t reset);

#ifdef __cplusplus
}
#endif
#endif
