/*
 * fsb_oracle.h -- CPU ORACLE for the FundStabABM simulation step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (fundstababm.jl_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * It restates, in plain C and IEEE FP64, the algorithm of the reference
 * /root/reference/src/functions.jl:13-491 (init :13-145,417-430; step :147-415; loop/log
 * :446-491), with the dense column-major arrays of src/types.jl:15-50.
 *
 * PARITY PINNING: the RNG-free golden vectors of test/functions_test.jl are
 * reproduced by tests/test_oracle_goldens.py.  What is NOT pinned (no Julia in this image):
 * Julia's MersenneTwister streams (replaced by Philox4x32-10 counters or by explicit tapes),
 * Distributions.jl's Categorical sampler (restated from its published algorithm), and the
 * summation order of BLAS/Base.sum (replaced by one canonical order, see below).
 *
 * Canonical FP order (shared by oracle and CUDA engine so that results are bit-identical):
 *   dotw(x,y,n)     : 8 partial fma-chains from +0.0, chain j owning elements 16i+2j and 16i+2j+1
 *                     (ascending), combined by the xor-butterfly tree 4,2,1 -- the order a group
 *                     of 8 GPU lanes produces when it reads a row with 128-bit loads and shuffles
 *   sumw(x,n)       : same tree with a_l += x_e
 *   sumprodw(x,y,n) : same tree with a_l += fl(x_e*y_e)     (reference materialises x.*y first)
 *   sums over orders (column sums)      : sequential in order-row order, from +0.0
 *   sum over a fund's investors (stakes, liquidate!): the fund's investors in ascending id dealt
 *                     round-robin to 32 partial sums from +0.0, combined by the tree 16,8,4,2,1
 */
#ifndef FSB_ORACLE_H
#define FSB_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors Params.default, /root/reference/src/params.jl:4-31 (ranges flattened) */
typedef struct {
    int64_t bigm, bign, bigt, bigk;
    double marketstartval, drift, marketvol;
    const double *impact_values; int64_t n_impact;   /* collect(impactrange) */
    double betamean, betastd, stockstartval;
    const double *vol_values; int64_t n_vol;         /* collect(stockvolrange) */
    double reviewprobability;
    int64_t perf_lo, perf_hi;                        /* perfwindow */
    int64_t cap_lo, cap_hi;                          /* invcaprange */
    double thresholdmean, thresholdstd;
    int64_t psize_lo, psize_hi;                      /* portfsizerange */
} orc_params;

enum { ORC_SEL_BEST = 0, ORC_SEL_GOODENOUGH = 1, ORC_SEL_PROBABILISTIC = 2, ORC_SEL_RANDOM = 3 };

/* flags for orc_modelrun */
enum {
    ORC_FAITHFUL_HISTORY = 1, /* fundrevalue! recomputes the whole K x T history (reference's work) */
    ORC_NO_VOLUME_QUIRK  = 2  /* record sell volume once per step instead of Q3's 1-3 times */
};

/* array selectors for orc_ptr */
enum {
    ORC_A_MARKET = 0, ORC_A_STOCK_VALUE, ORC_A_BETA, ORC_A_VOL, ORC_A_IMPACT, ORC_A_VOLUME,
    ORC_A_INV_ASSETS, ORC_A_THRESHOLD, ORC_A_HOLDINGS, ORC_A_STAKES, ORC_A_FUND_VALUE,
    ORC_A_TRACE_NAV_OPEN, ORC_A_TRACE_P_SELL, ORC_A_TRACE_P_RESP, ORC_A_NAV_CLOSE
};

/* trace / tape event kinds */
enum {
    ORC_EV_REVIEWER = 1, ORC_EV_DIVEST = 2, ORC_EV_CASHOUT = 3, ORC_EV_FLOOR = 4,
    ORC_EV_SPAWN = 5, ORC_EV_SPAWN_N = 6, ORC_EV_CHOICE = 7, ORC_EV_PICK = 8
};

/* tape sites for sparse (event-driven) draws, see orc_set_event_tape */
enum {
    ORC_TAPE_ANCHOR = 1, ORC_TAPE_RPSIZE = 2, ORC_TAPE_RPPICK = 3, ORC_TAPE_CHOICE_U = 4,
    ORC_TAPE_CHOICE_IDX = 5
};

typedef struct { int32_t t, kind, a, b; double v; } orc_trace_rec;
typedef struct { uint64_t key; double value; } orc_tape_rec;

/* error codes (also used as the replication error flag) */
enum {
    ORC_OK = 0, ORC_ERR_Q6_NO_ANCHOR = 1, ORC_ERR_TAPE_MISSING = 2, ORC_ERR_BAD_SELECTOR = 3,
    ORC_ERR_BAD_STATE = 4, ORC_ERR_TAPE_INVALID = 5
};

typedef struct orc_world orc_world;

orc_world *orc_create(const orc_params *p);
void orc_destroy(orc_world *w);
double *orc_ptr(orc_world *w, int which);
int64_t *orc_horizon(orc_world *w);

/* draws: Philox4x32-10 keyed (seed) with counter (block, a, site|b<<8, rep) ... */
void orc_set_rng(orc_world *w, uint64_t seed, uint32_t rep);
/* ... optionally overridden by dense tapes (NULL = keep Philox); column-major, t is 1-based:
 * z_mkt[T], z_stock[M x T], u_review[N x T] */
void orc_set_dense_tape(orc_world *w, const double *z_mkt, const double *z_stock, const double *u_review);
/* ... and by a sparse event tape sorted by key = orc_tape_key(site,t,a,b) */
void orc_set_event_tape(orc_world *w, const orc_tape_rec *recs, int64_t n);
uint64_t orc_tape_key(int site, int64_t t, int64_t a, int64_t b);
/* one day's exogenous draws of replications rep0 .. rep0+nreps-1 in the engine's host-tape layout */
void orc_fill_step_tape(uint64_t seed, uint32_t rep0, int64_t nreps, int64_t t, int64_t M, int64_t N,
                        double *z_mkt, double *z_stock, double *u_review);

void orc_enable_trace(orc_world *w, int on);
int64_t orc_trace_count(orc_world *w);
const orc_trace_rec *orc_trace(orc_world *w);

/* Func.initialise, functions.jl:417-430 */
int orc_initialise(orc_world *w);
/* Func.modelrun, functions.jl:446-491; t is 1-based and inclusive; builds the log on first call */
int orc_modelrun(orc_world *w, int64_t t_first, int64_t t_last, int selector, int flags);
/* funds.value as the reference leaves it (needed after lazy runs) */
void orc_finalise_fund_value(orc_world *w, int64_t t_now);

int64_t orc_log_rows(orc_world *w);
/* columns: 0 liquidated 1 replacedby 2 initialsize 3 liquidity 4 failtime 5 ninvestors 6 nstocks */
const double *orc_log_col(orc_world *w, int col);
int64_t orc_floor_hits(orc_world *w);

/* ---- per-function entry points (for the golden-vector tests); indices 0-based ---- */
void orc_fundcapitalinit(orc_world *w);
void orc_fundstakeinit(orc_world *w);
void orc_fundvalinit(orc_world *w, int64_t horizon);
void orc_fundrevalue(orc_world *w, int64_t k);               /* full history of fund k */
int64_t orc_perfreview(orc_world *w, int64_t t, const int64_t *reviewers, int64_t nrev,
                       int64_t *div_inv, int64_t *div_fund);
/* liquidate! with price column `col` (1-based); sale values out as R x M row-major */
void orc_liquidate(orc_world *w, int64_t col, const int64_t *div_inv, const int64_t *div_fund,
                   int64_t ndiv, double *sale_values);
void orc_marketmake(orc_world *w, int64_t t, const double *order_values, int64_t norders);
void orc_executeorder_sell(orc_world *w, int64_t t, const double *order_values, int64_t norders,
                           double *cash_out);
void orc_executeorder_buy(orc_world *w, int64_t t, const double *order_values, int64_t norders,
                          double *shares_out);
void orc_disburse_cash(orc_world *w, const int64_t *div_inv, const int64_t *div_fund, int64_t ndiv,
                       const double *cash);
void orc_disburse_shares(orc_world *w, const int64_t *funds, int64_t norders, const double *shares);
/* respawn! for given anchors/picks is exercised through the event tape in orc_respawn */
int orc_respawn(orc_world *w, int64_t t, double *buy_values, int64_t *buy_funds, int64_t *nbuy);
int orc_reinvest(orc_world *w, int64_t t, int selector, double *buy_values, int64_t *buy_funds,
                 int64_t *nbuy);

/* Func.calc_stylisedfacts, functions.jl:604-648 (array outputs; parity unpinned, see fsb_oracle.c).
 * Inputs column-major as Julia stores them: marketval[T], stocksval[M x T], stocksvolume[M x T].
 * Outputs: stock_pacf / stock_volacluster [M x maxlag] column-major, market_* [maxlag],
 * kurtoses / lossgain / corrs [M]. */
int orc_stylisedfacts(const double *marketval, const double *stocksval, const double *stocksvolume,
                      int64_t M, int64_t T, int64_t W, int64_t maxlag, double pctile,
                      double *stock_pacf, double *market_pacf, double *stock_volacluster,
                      double *market_volacluster, double *kurtoses, double *lossgain, double *corrs);

/* ---- RNG building blocks exposed for cross-checks against the CUDA engine ---- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_det_log(double x);
double orc_det_norminv(double u);
double orc_draw_normal(uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b, uint32_t idx);
double orc_draw_uniform(uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b, uint32_t idx);
uint32_t orc_draw_discrete(uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b,
                           uint32_t idx, uint32_t n);
double orc_dotw(const double *x, int64_t sx, const double *y, int64_t sy, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
