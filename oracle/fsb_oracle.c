/*
 * fsb_oracle.c -- CPU ORACLE (test infrastructure; see fsb_oracle.h for the rules of use).
 *
 * A plain-C, IEEE-FP64 restatement of /root/reference/src/functions.jl:13-491.
 * Every function cites the reference lines it follows.  Arrays are dense and column-major,
 * exactly as Julia stores the structs of /root/reference/src/types.jl:15-50:
 *   A[i,j] (1-based) lives at A[(i-1) + rows*(j-1)].
 * Internally investor / fund / stock ids are 0-based; time t and price columns are 1-based.
 *
 * Build: gcc -O2 -mfma -ffp-contract=off (no -ffast-math): every fused multiply-add in this
 * file is an explicit fma(), every other operation is a separately rounded IEEE operation.
 */
#include "fsb_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Counter-based draws: Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09 constants).
 * counter = (block, a, site | b<<8, rep), key = (seed_lo, seed_hi).
 * ---------------------------------------------------------------------------------------- */
enum {
    SITE_INIT_MKT = 1, SITE_INIT_BETA = 2, SITE_INIT_VOL = 3, SITE_INIT_STOCK = 4,
    SITE_INIT_IMPACT = 5, SITE_INIT_HORIZON = 6, SITE_INIT_THRESH = 7, SITE_INIT_CAPITAL = 8,
    SITE_INIT_FUNDPICK = 9, SITE_INIT_PSIZE = 10, SITE_INIT_PPICK = 11,
    SITE_MKT = 16, SITE_STOCK = 17, SITE_REVIEW = 18, SITE_SHUFFLE = 19, SITE_RPSIZE = 20,
    SITE_RPPICK = 21, SITE_CHOICE = 22
};

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void philox_block(uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b,
                         uint32_t block, uint32_t out[4])
{
    uint32_t ctr[4] = { block, a, site | (b << 8), rep };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    orc_philox4x32_10(ctr, key, out);
}

/* 52 random bits -> u in (0,1): (x + 0.5) * 2^-52, exactly representable */
static double u52(uint32_t hi, uint32_t lo)
{
    uint64_t x = (((uint64_t)hi << 32) | lo) >> 12;
    return ((double)x + 0.5) * 0x1.0p-52;
}

/* log(x) for x in (0,1] from basic IEEE operations only (bit-reproducible on any IEEE machine):
 * x = 2^e * m, m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1); log m = 2 s (1 + s^2/3 + s^4/5 + ...) */
double orc_det_log(double x)
{
    union { double d; uint64_t u; } v; v.d = x;
    int64_t e = (int64_t)((v.u >> 52) & 0x7ff) - 1023;
    v.u = (v.u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL; /* m in [1,2) */
    double m = v.d;
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double f = m - 1.0;
    double s = f / (m + 1.0);
    double z = s * s;
    /* Horner over 1/(2k+1), k = 13 .. 1 (|s| <= 0.1716 -> truncation < 1e-21) */
    double p = 1.0 / 27.0;
    p = fma(p, z, 1.0 / 25.0);
    p = fma(p, z, 1.0 / 23.0);
    p = fma(p, z, 1.0 / 21.0);
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    p = fma(p, z, 1.0);
    double lm = 2.0 * s * p;
    /* e*ln2 split hi/lo so that e*hi is exact */
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    double de = (double)e;
    return fma(de, ln2_hi, fma(de, ln2_lo, lm));
}

/* Inverse normal CDF, Wichura's AS241 PPND16 (Appl. Statist. 37 (1988) 477-484), with
 * orc_det_log in place of libm's log; Horner steps are explicit fma. */
double orc_det_norminv(double u)
{
    double q = u - 0.5, r, num, den, val;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        num = 2.5090809287301226727e+3;
        num = fma(num, r, 3.3430575583588128105e+4);
        num = fma(num, r, 6.7265770927008700853e+4);
        num = fma(num, r, 4.5921953931549871457e+4);
        num = fma(num, r, 1.3731693765509461125e+4);
        num = fma(num, r, 1.9715909503065514427e+3);
        num = fma(num, r, 1.3314166789178437745e+2);
        num = fma(num, r, 3.3871328727963666080e+0);
        den = 5.2264952788528545610e+3;
        den = fma(den, r, 2.8729085735721942674e+4);
        den = fma(den, r, 3.9307895800092710610e+4);
        den = fma(den, r, 2.1213794301586595867e+4);
        den = fma(den, r, 5.3941960214247511077e+3);
        den = fma(den, r, 6.8718700749205790830e+2);
        den = fma(den, r, 4.2313330701600911252e+1);
        den = fma(den, r, 1.0);
        return q * num / den;
    }
    r = (q < 0.0) ? u : 1.0 - u;
    r = sqrt(-orc_det_log(r));
    if (r <= 5.0) {
        r = r - 1.6;
        num = 7.74545014278341407640e-4;
        num = fma(num, r, 2.27238449892691845833e-2);
        num = fma(num, r, 2.41780725177450611770e-1);
        num = fma(num, r, 1.27045825245236838258e+0);
        num = fma(num, r, 3.64784832476320460504e+0);
        num = fma(num, r, 5.76949722146069140550e+0);
        num = fma(num, r, 4.63033784615654529590e+0);
        num = fma(num, r, 1.42343711074968357734e+0);
        den = 1.05075007164441684324e-9;
        den = fma(den, r, 5.47593808499534494600e-4);
        den = fma(den, r, 1.51986665636164571966e-2);
        den = fma(den, r, 1.48103976427480074590e-1);
        den = fma(den, r, 6.89767334985100004550e-1);
        den = fma(den, r, 1.67638483018380384940e+0);
        den = fma(den, r, 2.05319162663775882187e+0);
        den = fma(den, r, 1.0);
    } else {
        r = r - 5.0;
        num = 2.01033439929228813265e-7;
        num = fma(num, r, 2.71155556874348757815e-5);
        num = fma(num, r, 1.24266094738807843860e-3);
        num = fma(num, r, 2.65321895265761230930e-2);
        num = fma(num, r, 2.96560571828504891230e-1);
        num = fma(num, r, 1.78482653991729133580e+0);
        num = fma(num, r, 5.46378491116411436990e+0);
        num = fma(num, r, 6.65790464350110377720e+0);
        den = 2.04426310338993978564e-15;
        den = fma(den, r, 1.42151175831644588870e-7);
        den = fma(den, r, 1.84631831751005468180e-5);
        den = fma(den, r, 7.86869131145613259100e-4);
        den = fma(den, r, 1.48753612908506148525e-2);
        den = fma(den, r, 1.36929880922735805310e-1);
        den = fma(den, r, 5.99832206555887937690e-1);
        den = fma(den, r, 1.0);
    }
    val = num / den;
    return (q < 0.0) ? -val : val;
}

double orc_draw_uniform(uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b, uint32_t idx)
{
    uint32_t w[4];
    philox_block(seed, rep, site, a, b, idx >> 1, w);
    return (idx & 1) ? u52(w[2], w[3]) : u52(w[0], w[1]);
}

double orc_draw_normal(uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b, uint32_t idx)
{
    return orc_det_norminv(orc_draw_uniform(seed, rep, site, a, b, idx));
}

/* unbiased-enough integer in [0,n): (word * n) >> 32 */
uint32_t orc_draw_discrete(uint64_t seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b,
                           uint32_t idx, uint32_t n)
{
    uint32_t w[4];
    philox_block(seed, rep, site, a, b, idx >> 2, w);
    return (uint32_t)(((uint64_t)w[idx & 3] * (uint64_t)n) >> 32);
}

/* ------------------------------------------------------------------------------------------
 * canonical reductions
 * ---------------------------------------------------------------------------------------- */
/* 8 partial chains (chain j owns elements 16i+2j, 16i+2j+1, ascending, from +0.0) followed by the
 * xor-butterfly tree 4,2,1 -- the order a group of 8 GPU lanes reading a row with 128-bit loads
 * produces (the engine maps 4 such groups, i.e. 4 rows, onto one warp) */
#define ORC_CHAINS 8
static double chain_tree(double a[ORC_CHAINS])
{
    for (int s = ORC_CHAINS / 2; s >= 1; s >>= 1)
        for (int l = 0; l < s; ++l) a[l] = a[l] + a[l + s];
    return a[0];
}
double orc_dotw(const double *x, int64_t sx, const double *y, int64_t sy, int64_t n)
{
    double a[ORC_CHAINS] = { 0.0 };
    for (int64_t e = 0; e < n; ++e) { int l = (int)((e & (2 * ORC_CHAINS - 1)) >> 1); a[l] = fma(x[e * sx], y[e * sy], a[l]); }
    return chain_tree(a);
}
static double sumw(const double *x, int64_t sx, int64_t n)
{
    double a[ORC_CHAINS] = { 0.0 };
    for (int64_t e = 0; e < n; ++e) { int l = (int)((e & (2 * ORC_CHAINS - 1)) >> 1); a[l] = a[l] + x[e * sx]; }
    return chain_tree(a);
}
static double sumprodw(const double *x, int64_t sx, const double *y, int64_t sy, int64_t n)
{
    double a[ORC_CHAINS] = { 0.0 };
    for (int64_t e = 0; e < n; ++e) { int l = (int)((e & (2 * ORC_CHAINS - 1)) >> 1); double p = x[e * sx] * y[e * sy]; a[l] = a[l] + p; }
    return chain_tree(a);
}

/* Julia's isless for Float64: total order, -0.0 < +0.0, NaN last */
static int jl_isless(double a, double b)
{
    if (isnan(a)) return 0;
    if (isnan(b)) return 1;
    if (a < b) return 1;
    if (a == b) return signbit(a) && !signbit(b);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * world
 * ---------------------------------------------------------------------------------------- */
struct orc_world {
    orc_params p;
    double *impact_values, *vol_values;
    int64_t K, M, N, T, W;
    /* types.jl:15-50 */
    double *market;      /* T */
    double *stock_value; /* M x T */
    double *beta, *vol, *impact; /* M */
    double *volume;      /* M x T */
    double *inv_assets;  /* N x (K+1) */
    int64_t *horizon;    /* N */
    double *threshold;   /* N */
    double *holdings;    /* K x M */
    double *stakes;      /* K x N */
    double *fund_value;  /* K x T */
    double *nav_close;   /* K: what funds.value[:,t] holds at the end of step t */
    /* liquidation log, functions.jl:452-456 */
    int log_built; int64_t log_rows, log_cap; double *log_cols[7];
    /* draws */
    uint64_t seed; uint32_t rep;
    const double *z_mkt, *z_stock, *u_review;
    const orc_tape_rec *tape; int64_t n_tape; int tape_mode;
    /* trace */
    int trace_on; orc_trace_rec *trace; int64_t n_trace, cap_trace;
    double *tr_nav_open, *tr_p_sell, *tr_p_resp; /* K x T, M x T, M x T */
    int64_t floor_hits;
    int faithful;
    int err;
};

#define P_(w, m, t) ((w)->stock_value[(m) + (w)->M * ((t)-1)])
#define H_(w, k, m) ((w)->holdings[(k) + (w)->K * (m)])
#define S_(w, k, n) ((w)->stakes[(k) + (w)->K * (n)])
#define A_(w, n, c) ((w)->inv_assets[(n) + (w)->N * (c)])
#define V_(w, k, t) ((w)->fund_value[(k) + (w)->K * ((t)-1)])

static void *zalloc(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) abort(); return p; }

orc_world *orc_create(const orc_params *p)
{
    orc_world *w = (orc_world *)zalloc(sizeof *w);
    w->p = *p;
    w->impact_values = (double *)zalloc(sizeof(double) * p->n_impact);
    memcpy(w->impact_values, p->impact_values, sizeof(double) * p->n_impact);
    w->vol_values = (double *)zalloc(sizeof(double) * p->n_vol);
    memcpy(w->vol_values, p->vol_values, sizeof(double) * p->n_vol);
    w->p.impact_values = w->impact_values; w->p.vol_values = w->vol_values;
    int64_t K = p->bigk, M = p->bigm, N = p->bign, T = p->bigt;
    w->K = K; w->M = M; w->N = N; w->T = T; w->W = p->perf_hi;
    w->market = (double *)zalloc(8 * T);
    w->stock_value = (double *)zalloc(8 * M * T);
    w->beta = (double *)zalloc(8 * M); w->vol = (double *)zalloc(8 * M); w->impact = (double *)zalloc(8 * M);
    w->volume = (double *)zalloc(8 * M * T);
    w->inv_assets = (double *)zalloc(8 * N * (K + 1));
    w->horizon = (int64_t *)zalloc(8 * N);
    w->threshold = (double *)zalloc(8 * N);
    w->holdings = (double *)zalloc(8 * K * M);
    w->stakes = (double *)zalloc(8 * K * N);
    w->fund_value = (double *)zalloc(8 * K * T);
    w->nav_close = (double *)zalloc(8 * K);
    return w;
}

void orc_destroy(orc_world *w)
{
    if (!w) return;
    free(w->impact_values); free(w->vol_values); free(w->market); free(w->stock_value);
    free(w->beta); free(w->vol); free(w->impact); free(w->volume); free(w->inv_assets);
    free(w->horizon); free(w->threshold); free(w->holdings); free(w->stakes); free(w->fund_value);
    free(w->nav_close);
    for (int c = 0; c < 7; ++c) free(w->log_cols[c]);
    free(w->trace); free(w->tr_nav_open); free(w->tr_p_sell); free(w->tr_p_resp);
    free(w);
}

double *orc_ptr(orc_world *w, int which)
{
    switch (which) {
    case ORC_A_MARKET: return w->market;
    case ORC_A_STOCK_VALUE: return w->stock_value;
    case ORC_A_BETA: return w->beta;
    case ORC_A_VOL: return w->vol;
    case ORC_A_IMPACT: return w->impact;
    case ORC_A_VOLUME: return w->volume;
    case ORC_A_INV_ASSETS: return w->inv_assets;
    case ORC_A_THRESHOLD: return w->threshold;
    case ORC_A_HOLDINGS: return w->holdings;
    case ORC_A_STAKES: return w->stakes;
    case ORC_A_FUND_VALUE: return w->fund_value;
    case ORC_A_TRACE_NAV_OPEN: return w->tr_nav_open;
    case ORC_A_TRACE_P_SELL: return w->tr_p_sell;
    case ORC_A_TRACE_P_RESP: return w->tr_p_resp;
    case ORC_A_NAV_CLOSE: return w->nav_close;
    }
    return 0;
}
int64_t *orc_horizon(orc_world *w) { return w->horizon; }

void orc_set_rng(orc_world *w, uint64_t seed, uint32_t rep) { w->seed = seed; w->rep = rep; }
void orc_set_dense_tape(orc_world *w, const double *z_mkt, const double *z_stock, const double *u_review)
{
    w->z_mkt = z_mkt; w->z_stock = z_stock; w->u_review = u_review;
}
void orc_set_event_tape(orc_world *w, const orc_tape_rec *recs, int64_t n)
{
    w->tape = recs; w->n_tape = n; w->tape_mode = (recs != 0);
}
uint64_t orc_tape_key(int site, int64_t t, int64_t a, int64_t b)
{
    return ((uint64_t)(site & 0xff) << 56) | ((uint64_t)(t & 0xffffff) << 32) |
           ((uint64_t)(a & 0xffff) << 16) | (uint64_t)(b & 0xffff);
}
static int tape_lookup(orc_world *w, int site, int64_t t, int64_t a, int64_t b, double *out)
{
    uint64_t key = orc_tape_key(site, t, a, b);
    int64_t lo = 0, hi = w->n_tape - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        if (w->tape[mid].key == key) { *out = w->tape[mid].value; return 1; }
        if (w->tape[mid].key < key) lo = mid + 1; else hi = mid - 1;
    }
    return 0;
}

void orc_enable_trace(orc_world *w, int on)
{
    w->trace_on = on;
    if (on && !w->tr_nav_open) {
        w->tr_nav_open = (double *)zalloc(8 * w->K * w->T);
        w->tr_p_sell = (double *)zalloc(8 * w->M * w->T);
        w->tr_p_resp = (double *)zalloc(8 * w->M * w->T);
    }
}
int64_t orc_trace_count(orc_world *w) { return w->n_trace; }
const orc_trace_rec *orc_trace(orc_world *w) { return w->trace; }
static void emit(orc_world *w, int64_t t, int kind, int64_t a, int64_t b, double v)
{
    if (!w->trace_on) return;
    if (w->n_trace == w->cap_trace) {
        w->cap_trace = w->cap_trace ? 2 * w->cap_trace : 1024;
        w->trace = (orc_trace_rec *)realloc(w->trace, sizeof(orc_trace_rec) * w->cap_trace);
        if (!w->trace) abort();
    }
    orc_trace_rec *r = &w->trace[w->n_trace++];
    r->t = (int32_t)t; r->kind = kind; r->a = (int32_t)a; r->b = (int32_t)b; r->v = v;
}

/* draw helpers ---------------------------------------------------------------------------- */
static double draw_z_mkt(orc_world *w, int64_t t)
{
    if (w->z_mkt) return w->z_mkt[t - 1];
    return orc_draw_normal(w->seed, w->rep, SITE_MKT, (uint32_t)t, 0, 0);
}
static double draw_z_stock(orc_world *w, int64_t t, int64_t m)
{
    if (w->z_stock) return w->z_stock[m + w->M * (t - 1)];
    return orc_draw_normal(w->seed, w->rep, SITE_STOCK, (uint32_t)t, 0, (uint32_t)m);
}
static double draw_u_review(orc_world *w, int64_t t, int64_t n)
{
    if (w->u_review) return w->u_review[n + w->N * (t - 1)];
    return orc_draw_uniform(w->seed, w->rep, SITE_REVIEW, (uint32_t)t, 0, (uint32_t)n);
}

/* One day's exogenous draws of a batch of replications as the dense "host tape" of the engine's batched equivalence mode
 * (fsb_run_step_hosttape): z_mkt[nreps], z_stock[nreps][M], u_review[nreps][N] -- exactly the values the draw helpers above
 * produce for (seed, rep0 + i, t), i.e. what a recorder wrapped around functions.jl:15, :32, :149 would capture. */
void orc_fill_step_tape(uint64_t seed, uint32_t rep0, int64_t nreps, int64_t t, int64_t M, int64_t N,
                        double *z_mkt, double *z_stock, double *u_review)
{
    for (int64_t i = 0; i < nreps; ++i) {
        const uint32_t rep = rep0 + (uint32_t)i;
        z_mkt[i] = orc_draw_normal(seed, rep, SITE_MKT, (uint32_t)t, 0, 0);
        for (int64_t m = 0; m < M; ++m) z_stock[i * M + m] = orc_draw_normal(seed, rep, SITE_STOCK, (uint32_t)t, 0, (uint32_t)m);
        for (int64_t n = 0; n < N; ++n) u_review[i * N + n] = orc_draw_uniform(seed, rep, SITE_REVIEW, (uint32_t)t, 0, (uint32_t)n);
    }
}

/* ------------------------------------------------------------------------------------------
 * price functions
 * ---------------------------------------------------------------------------------------- */
/* marketmove!, functions.jl:13-17:  mkt[t] = mkt[t-1]*(1 + drift + marketvol*z) */
static void marketmove(orc_world *w, int64_t t, double z)
{
    double g = (1.0 + w->p.drift) + w->p.marketvol * z;
    w->market[t - 1] = w->market[t - 2] * g;
}

/* stockmove!, functions.jl:29-34 */
static void stockmove(orc_world *w, int64_t t, const double *z)
{
    double mktreturn = (w->market[t - 1] - w->market[t - 2]) / w->market[t - 2];
    for (int64_t m = 0; m < w->M; ++m) {
        double prev = P_(w, m, t - 1);
        double a = (1.0 + mktreturn * w->beta[m]) * prev;
        double b = (w->vol[m] * z[m]) * prev;
        P_(w, m, t) = a + b;
    }
}

/* fundrevalue!, functions.jl:266-271: whole history of fund k with current holdings */
void orc_fundrevalue(orc_world *w, int64_t k)
{
    for (int64_t t = 1; t <= w->T; ++t)
        V_(w, k, t) = orc_dotw(&H_(w, k, 0), w->K, &P_(w, 0, t), 1, w->M);
}

/* NAV of fund k against price column col with CURRENT holdings (Q2: constant-composition
 * back-cast).  Equals funds.value[k,col] whenever the reference reads it (functions.jl:164,377). */
static double nav(orc_world *w, int64_t k, int64_t col)
{
    if (w->faithful) return V_(w, k, col);
    return orc_dotw(&H_(w, k, 0), w->K, &P_(w, 0, col), 1, w->M);
}

/* ------------------------------------------------------------------------------------------
 * initialisation, functions.jl:19-145, 417-430 (draw order of SURVEY 3.2; here every draw is
 * addressed by (site, index) instead of by stream position)
 * ---------------------------------------------------------------------------------------- */
/* fundcapitalinit!, functions.jl:104-107 */
void orc_fundcapitalinit(orc_world *w)
{
    for (int64_t k = 0; k < w->K; ++k) {
        double s = 0.0;
        for (int64_t n = 0; n < w->N; ++n) s = s + A_(w, n, k);
        V_(w, k, 1) = s;
    }
}
/* fundstakeinit!, functions.jl:111-117 */
void orc_fundstakeinit(orc_world *w)
{
    for (int64_t k = 0; k < w->K; ++k) {
        double s = 0.0;
        for (int64_t n = 0; n < w->N; ++n) s = s + A_(w, n, k);
        for (int64_t n = 0; n < w->N; ++n) S_(w, k, n) = A_(w, n, k) / s;
    }
}
/* fundvalinit!, functions.jl:135-145 */
void orc_fundvalinit(orc_world *w, int64_t horizon)
{
    for (int64_t t = 2; t <= horizon + 1; ++t)
        for (int64_t k = 0; k < w->K; ++k)
            V_(w, k, t) = orc_dotw(&H_(w, k, 0), w->K, &P_(w, 0, t), 1, w->M);
}

int orc_initialise(orc_world *w)
{
    const orc_params *p = &w->p;
    int64_t K = w->K, M = w->M, N = w->N, W = w->W;
    uint64_t sd = w->seed; uint32_t rp = w->rep;
    /* marketinit!, :19-27 */
    w->market[0] = p->marketstartval;
    for (int64_t t = 2; t <= W + 1; ++t)
        marketmove(w, t, orc_draw_normal(sd, rp, SITE_INIT_MKT, 0, 0, (uint32_t)t));
    /* betainit!, :48-52 */
    for (int64_t m = 0; m < M; ++m)
        w->beta[m] = orc_draw_normal(sd, rp, SITE_INIT_BETA, 0, 0, (uint32_t)m) * p->betastd + p->betamean;
    /* stockvolinit!, :55-59 */
    for (int64_t m = 0; m < M; ++m)
        w->vol[m] = p->vol_values[orc_draw_discrete(sd, rp, SITE_INIT_VOL, 0, 0, (uint32_t)m, (uint32_t)p->n_vol)];
    /* stockvalueinit!, :61-70 */
    double *z = (double *)zalloc(8 * M);
    for (int64_t m = 0; m < M; ++m) P_(w, m, 1) = p->stockstartval;
    for (int64_t t = 2; t <= W + 1; ++t) {
        for (int64_t m = 0; m < M; ++m) z[m] = orc_draw_normal(sd, rp, SITE_INIT_STOCK, (uint32_t)t, 0, (uint32_t)m);
        stockmove(w, t, z);
    }
    free(z);
    /* stockimpactinit!, :72-76 */
    for (int64_t m = 0; m < M; ++m)
        w->impact[m] = p->impact_values[orc_draw_discrete(sd, rp, SITE_INIT_IMPACT, 0, 0, (uint32_t)m, (uint32_t)p->n_impact)];
    /* invhorizoninit!, :78-82 */
    for (int64_t n = 0; n < N; ++n)
        w->horizon[n] = p->perf_lo + orc_draw_discrete(sd, rp, SITE_INIT_HORIZON, 0, 0, (uint32_t)n, (uint32_t)(p->perf_hi - p->perf_lo + 1));
    /* invthreshinit!, :84-88 */
    for (int64_t n = 0; n < N; ++n)
        w->threshold[n] = orc_draw_normal(sd, rp, SITE_INIT_THRESH, 0, 0, (uint32_t)n) * p->thresholdstd + p->thresholdmean;
    /* invassetinit!, :91-102 */
    memset(w->inv_assets, 0, 8 * N * (K + 1));
    for (int64_t n = 0; n < N; ++n) {
        double cap = (double)(p->cap_lo + orc_draw_discrete(sd, rp, SITE_INIT_CAPITAL, 0, 0, (uint32_t)n, (uint32_t)(p->cap_hi - p->cap_lo + 1)));
        int64_t f = (n < K) ? n : (int64_t)orc_draw_discrete(sd, rp, SITE_INIT_FUNDPICK, 0, 0, (uint32_t)n, (uint32_t)K);
        A_(w, n, f) = cap;
    }
    orc_fundcapitalinit(w);
    orc_fundstakeinit(w);
    /* fundholdinit!, :119-133 (stockvals[stock] indexes column 1 of the M x T matrix) */
    memset(w->holdings, 0, 8 * K * M);
    for (int64_t k = 0; k < K; ++k) {
        int64_t nk = p->psize_lo + orc_draw_discrete(sd, rp, SITE_INIT_PSIZE, 0, 0, (uint32_t)k, (uint32_t)(p->psize_hi - p->psize_lo + 1));
        for (int64_t j = 0; j < nk; ++j) {
            int64_t s = orc_draw_discrete(sd, rp, SITE_INIT_PPICK, (uint32_t)k, 0, (uint32_t)j, (uint32_t)M);
            H_(w, k, s) = H_(w, k, s) + (V_(w, k, 1) / (double)nk) / P_(w, s, 1);
        }
    }
    orc_fundvalinit(w, W);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * agent behaviours
 * ---------------------------------------------------------------------------------------- */
/* perfreview, functions.jl:157-176 (Q1: denominator is column `horizon`) -- matrix version */
int64_t orc_perfreview(orc_world *w, int64_t t, const int64_t *reviewers, int64_t nrev,
                       int64_t *div_inv, int64_t *div_fund)
{
    int64_t nd = 0;
    int save = w->faithful; w->faithful = 1;
    for (int64_t i = 0; i < nrev; ++i) {
        int64_t rev = reviewers[i], h = w->horizon[rev];
        for (int64_t f = 0; f < w->K; ++f) {
            if (!(A_(w, rev, f) > 0.0)) continue;
            double valchange = (nav(w, f, t) - nav(w, f, t - h)) / nav(w, f, h);
            if (valchange < w->threshold[rev]) { div_inv[nd] = rev; div_fund[nd] = f; ++nd; }
        }
    }
    w->faithful = save;
    return nd;
}

/* liquidate!, functions.jl:183-201; sale_values is R x M row-major */
void orc_liquidate(orc_world *w, int64_t col, const int64_t *div_inv, const int64_t *div_fund,
                   int64_t ndiv, double *sale_values)
{
    int64_t M = w->M, N = w->N;
    for (int64_t r = 0; r < ndiv; ++r) {
        int64_t inv = div_inv[r], f = div_fund[r];
        double s = S_(w, f, inv);
        for (int64_t m = 0; m < M; ++m) sale_values[r * M + m] = (-s) * (H_(w, f, m) * P_(w, m, col));
        double remainder = 1.0 - s;
        for (int64_t m = 0; m < M; ++m) H_(w, f, m) = H_(w, f, m) * remainder;
        S_(w, f, inv) = 0.0;
        /* sum(stakes[fund,:]): canonical order = the fund's investors (non-zero position in the fund,
         * ascending id, the leaver included with its zeroed stake) dealt round-robin to 32 partial sums,
         * combined by the butterfly tree 16,8,4,2,1 */
        double part[32] = { 0.0 };
        int64_t idx = 0;
        for (int64_t n = 0; n < N; ++n)
            if (A_(w, n, f) != 0.0) { part[idx & 31] = part[idx & 31] + S_(w, f, n); ++idx; }
        for (int sft = 16; sft >= 1; sft >>= 1)
            for (int l = 0; l < sft; ++l) part[l] = part[l] + part[l + sft];
        double tot = part[0];
        if (tot > 0.0)
            for (int64_t n = 0; n < N; ++n) S_(w, f, n) = S_(w, f, n) / tot;
    }
}

/* marketmake!, functions.jl:205-211 (Q10 floor); order_values R x M row-major */
static int64_t marketmake_phase(orc_world *w, int64_t t, const double *order_values, int64_t norders, int phase)
{
    int64_t hits = 0;
    for (int64_t m = 0; m < w->M; ++m) {
        double net = 0.0;
        for (int64_t r = 0; r < norders; ++r) net = net + order_values[r * w->M + m];
        double netimpact = net * w->impact[m];
        double v = (1.0 + netimpact) * P_(w, m, t);
        if (v < 0.000000001) { v = 0.000000001; ++hits; emit(w, t, ORC_EV_FLOOR, m, phase, 0.0); }
        P_(w, m, t) = v;
    }
    w->floor_hits += hits;
    return hits;
}
void orc_marketmake(orc_world *w, int64_t t, const double *order_values, int64_t norders)
{
    marketmake_phase(w, t, order_values, norders, 0);
}

/* executeorder!(::SellMarketOrder), functions.jl:213-224 */
static void executeorder_sell_phase(orc_world *w, int64_t t, const double *order_values, int64_t norders,
                                    double *cash_out)
{
    int64_t M = w->M;
    double *old = (double *)zalloc(8 * M), *ratio = (double *)zalloc(8 * M);
    for (int64_t m = 0; m < M; ++m) old[m] = P_(w, m, t);
    marketmake_phase(w, t, order_values, norders, 0);
    for (int64_t m = 0; m < M; ++m) ratio[m] = P_(w, m, t) / old[m];
    for (int64_t r = 0; r < norders; ++r) cash_out[r] = -sumprodw(order_values + r * M, 1, ratio, 1, M);
    free(old); free(ratio);
}
void orc_executeorder_sell(orc_world *w, int64_t t, const double *order_values, int64_t norders, double *cash_out)
{
    executeorder_sell_phase(w, t, order_values, norders, cash_out);
}

/* executeorder!(::BuyMarketOrder), functions.jl:229-239; shares_out R x M row-major */
static void executeorder_buy_phase(orc_world *w, int64_t t, const double *order_values, int64_t norders,
                                   double *shares_out, int phase)
{
    marketmake_phase(w, t, order_values, norders, phase);
    for (int64_t r = 0; r < norders; ++r)
        for (int64_t m = 0; m < w->M; ++m) shares_out[r * w->M + m] = order_values[r * w->M + m] / P_(w, m, t);
}
void orc_executeorder_buy(orc_world *w, int64_t t, const double *order_values, int64_t norders, double *shares_out)
{
    executeorder_buy_phase(w, t, order_values, norders, shares_out, 1);
}

/* disburse!(invassets, divestments, cashout), functions.jl:243-252 (Q8: assignment) */
void orc_disburse_cash(orc_world *w, const int64_t *div_inv, const int64_t *div_fund, int64_t ndiv,
                       const double *cash)
{
    for (int64_t r = 0; r < ndiv; ++r) {
        A_(w, div_inv[r], div_fund[r]) = 0.0;
        A_(w, div_inv[r], w->K) = cash[r];
    }
}

/* disburse!(funds, sharesout, stockvals), functions.jl:256-263; t_now only for nav_close */
static void disburse_shares_at(orc_world *w, const int64_t *funds, int64_t norders, const double *shares, int64_t t_now)
{
    for (int64_t r = 0; r < norders; ++r) {
        int64_t f = funds[r];
        for (int64_t m = 0; m < w->M; ++m) H_(w, f, m) = H_(w, f, m) + shares[r * w->M + m];
        if (w->faithful) orc_fundrevalue(w, f);
        if (t_now > 0) w->nav_close[f] = orc_dotw(&H_(w, f, 0), w->K, &P_(w, 0, t_now), 1, w->M);
    }
}
void orc_disburse_shares(orc_world *w, const int64_t *funds, int64_t norders, const double *shares)
{
    int save = w->faithful; w->faithful = 1;
    disburse_shares_at(w, funds, norders, shares, 0);
    w->faithful = save;
}

/* ---- liquidation log ------------------------------------------------------------------- */
static void log_push(orc_world *w, double liquidated, double replacedby, double initialsize,
                     double liquidity, double failtime, double ninvestors, double nstocks)
{
    if (w->log_rows == w->log_cap) {
        w->log_cap = w->log_cap ? 2 * w->log_cap : 64;
        for (int c = 0; c < 7; ++c) {
            w->log_cols[c] = (double *)realloc(w->log_cols[c], 8 * w->log_cap);
            if (!w->log_cols[c]) abort();
        }
    }
    int64_t r = w->log_rows++;
    w->log_cols[0][r] = liquidated; w->log_cols[1][r] = replacedby; w->log_cols[2][r] = initialsize;
    w->log_cols[3][r] = liquidity; w->log_cols[4][r] = failtime; w->log_cols[5][r] = ninvestors;
    w->log_cols[6][r] = nstocks;
}
int64_t orc_log_rows(orc_world *w) { return w->log_rows; }
const double *orc_log_col(orc_world *w, int col) { return w->log_cols[col]; }
int64_t orc_floor_hits(orc_world *w) { return w->floor_hits; }

/* calc_liquidity, functions.jl:324-330 */
static double calc_liquidity(orc_world *w, int64_t k, int64_t col)
{
    int64_t M = w->M;
    double *wt = (double *)zalloc(8 * M);
    for (int64_t m = 0; m < M; ++m) wt[m] = H_(w, k, m) * P_(w, m, col);
    double s = sumw(wt, 1, M);
    for (int64_t m = 0; m < M; ++m) wt[m] = (wt[m] / s) * w->impact[m];
    double liq = sumw(wt, 1, M);
    free(wt);
    return liq;
}

/* log construction, functions.jl:448-456 (Q4: sized with K, not bigm) */
static void build_log(orc_world *w)
{
    for (int64_t k = 0; k < w->K; ++k) {
        int64_t ninv = 0, nst = 0;
        for (int64_t n = 0; n < w->N; ++n) ninv += (S_(w, k, n) > 0.0);
        for (int64_t m = 0; m < w->M; ++m) nst += (H_(w, k, m) > 0.0);
        log_push(w, 0, 0, V_(w, k, 1), calc_liquidity(w, k, 1), 0, (double)ninv, (double)nst);
    }
    w->log_built = 1;
}

/* respawn!, functions.jl:275-321 (+ fundholdinit! :119-133 on a copy of the egg's row).
 * buy_values: rows x M row-major. */
int orc_respawn(orc_world *w, int64_t t, double *buy_values, int64_t *buy_funds, int64_t *nbuy)
{
    int64_t K = w->K, M = w->M, N = w->N;
    const orc_params *p = &w->p;
    int64_t *spawn = (int64_t *)zalloc(8 * K), neggs = 0;
    for (int64_t k = 0; k < K; ++k)
        if (sumw(&H_(w, k, 0), K, M) == 0.0) spawn[neggs++] = k;       /* :279 */
    int64_t *pot = (int64_t *)zalloc(8 * N), npot = 0;
    for (int64_t n = 0; n < N; ++n)
        if (A_(w, n, K) > 0.0) pot[npot++] = n;                         /* :281 */
    *nbuy = 0;
    if (neggs > npot) { free(spawn); free(pot); return ORC_ERR_Q6_NO_ANCHOR; }   /* Q6 */
    double *units = (double *)zalloc(8 * M);
    int rc = ORC_OK;
    for (int64_t i = 0; i < neggs && rc == ORC_OK; ++i) {
        /* :282 shuffle(potentialinvs)[1:neggs] == partial Fisher-Yates, one draw per anchor */
        int64_t j;
        if (w->tape_mode) {
            double v;
            if (!tape_lookup(w, ORC_TAPE_ANCHOR, t, i, 0, &v)) { rc = ORC_ERR_TAPE_MISSING; break; }
            for (j = i; j < npot; ++j) if (pot[j] == (int64_t)v) break;
            if (j == npot) { rc = ORC_ERR_TAPE_INVALID; break; }
        } else {
            j = i + orc_draw_discrete(w->seed, w->rep, SITE_SHUFFLE, (uint32_t)t, 0, (uint32_t)i, (uint32_t)(npot - i));
        }
        int64_t tmp = pot[i]; pot[i] = pot[j]; pot[j] = tmp;
        int64_t inv = pot[i], egg = spawn[i];
        double capital = A_(w, inv, K);                                 /* :286 */
        A_(w, inv, egg) = capital; A_(w, inv, K) = 0.0;                 /* :288-289 */
        emit(w, t, ORC_EV_SPAWN, egg, inv, capital);
        /* fundholdinit! on the (all-zero) row copy, :123-129 */
        int64_t nk;
        if (w->tape_mode) {
            double v;
            if (!tape_lookup(w, ORC_TAPE_RPSIZE, t, i, 0, &v)) { rc = ORC_ERR_TAPE_MISSING; break; }
            nk = (int64_t)v;
        } else {
            nk = p->psize_lo + orc_draw_discrete(w->seed, w->rep, SITE_RPSIZE, (uint32_t)t, 0, (uint32_t)i, (uint32_t)(p->psize_hi - p->psize_lo + 1));
        }
        emit(w, t, ORC_EV_SPAWN_N, egg, nk, 0.0);
        for (int64_t m = 0; m < M; ++m) units[m] = H_(w, egg, m);
        for (int64_t jj = 0; jj < nk; ++jj) {
            int64_t s;
            if (w->tape_mode) {
                double v;
                if (!tape_lookup(w, ORC_TAPE_RPPICK, t, i, jj, &v)) { rc = ORC_ERR_TAPE_MISSING; break; }
                s = (int64_t)v;
            } else {
                s = orc_draw_discrete(w->seed, w->rep, SITE_RPPICK, (uint32_t)t, (uint32_t)i, (uint32_t)jj, (uint32_t)M);
            }
            emit(w, t, ORC_EV_PICK, egg, s, 0.0);
            units[s] = units[s] + (capital / (double)nk) / P_(w, s, t);
        }
        if (rc != ORC_OK) break;
        double *bo = buy_values + (*nbuy) * M;
        for (int64_t m = 0; m < M; ++m) bo[m] = units[m] * P_(w, m, t); /* :291-292 */
        buy_funds[*nbuy] = egg; ++*nbuy;
        S_(w, egg, inv) = 1.0;                                          /* :296 */
        /* log chain, :298-308 */
        int64_t idx_new = egg, idx_old = egg; double chk = 1.0;
        while (chk == 1.0) {
            idx_old = idx_new;
            chk = w->log_cols[0][idx_old];
            idx_new = (int64_t)w->log_cols[1][idx_old] - 1;
        }
        w->log_cols[0][idx_old] = 1.0;
        w->log_cols[1][idx_old] = (double)(w->log_rows + 1);
        w->log_cols[4][idx_old] = (double)t;
        /* new row, :311-318 */
        double fundsize = sumw(bo, 1, M);
        int64_t nst = 0;
        for (int64_t m = 0; m < M; ++m) nst += (bo[m] > 0.0);
        for (int64_t m = 0; m < M; ++m) units[m] = (bo[m] / fundsize) * w->impact[m];
        double liq = sumw(units, 1, M);
        log_push(w, 0, 0, fundsize, liq, 0, 1, (double)nst);
    }
    free(spawn); free(pot); free(units);
    return rc;
}

/* selectors, functions.jl:381-415 */
static int64_t sel_best(const double *rho, int64_t K)
{
    int64_t mi = 0; double m = rho[0];
    for (int64_t k = 1; k < K; ++k) {
        if (m != m) break;
        double a = rho[k];
        if (a != a || jl_isless(m, a)) { m = a; mi = k; }
    }
    return mi;
}

/* reinvest!, functions.jl:335-373 (Q7 stale fundval, Q12 NaN); buy_values rows x M row-major */
int orc_reinvest(orc_world *w, int64_t t, int selector, double *buy_values, int64_t *buy_funds,
                 int64_t *nbuy)
{
    int64_t K = w->K, M = w->M, N = w->N;
    *nbuy = 0;
    if (selector < 0 || selector > 3) return ORC_ERR_BAD_SELECTOR;
    double *rho = (double *)zalloc(8 * K), *vt = (double *)zalloc(8 * K);
    /* funds.value[:,t] as of the revaluation at :479 (holdings do not change inside this loop) */
    for (int64_t k = 0; k < K; ++k) vt[k] = nav(w, k, t);
    int rc = ORC_OK;
    for (int64_t reinv = 0; reinv < N && rc == ORC_OK; ++reinv) {
        if (!(A_(w, reinv, K) > 0.0)) continue;                         /* :339 */
        double injection = A_(w, reinv, K);
        int64_t h = w->horizon[reinv];
        for (int64_t k = 0; k < K; ++k) {                               /* :375-379 */
            double vh = nav(w, k, t - h);
            rho[k] = (vt[k] - vh) / vh;
        }
        int64_t choice = 0;
        double forced;
        if (w->tape_mode && tape_lookup(w, ORC_TAPE_CHOICE_IDX, t, reinv, 0, &forced)) {
            choice = (int64_t)forced;
        } else if (selector == ORC_SEL_BEST) {
            choice = sel_best(rho, K);
        } else {
            double u; uint32_t word;
            if (w->tape_mode) {
                if (!tape_lookup(w, ORC_TAPE_CHOICE_U, t, reinv, 0, &u)) { rc = ORC_ERR_TAPE_MISSING; break; }
                word = (uint32_t)(u * 4294967296.0);
            } else {
                uint32_t wd[4];
                philox_block(w->seed, w->rep, SITE_CHOICE, (uint32_t)t, 0, (uint32_t)reinv, wd);
                u = u52(wd[0], wd[1]); word = wd[2];
            }
            if (selector == ORC_SEL_RANDOM) {                           /* :412-415 */
                choice = (int64_t)(((uint64_t)word * (uint64_t)K) >> 32);
            } else if (selector == ORC_SEL_GOODENOUGH) {                /* :387-395 */
                int64_t nc = 0;
                for (int64_t k = 0; k < K; ++k) nc += (rho[k] > w->threshold[reinv]);
                if (nc == 0) choice = sel_best(rho, K);
                else {
                    int64_t pick = (int64_t)(((uint64_t)word * (uint64_t)nc) >> 32), c = 0;
                    for (int64_t k = 0; k < K; ++k)
                        if (rho[k] > w->threshold[reinv]) { if (c == pick) { choice = k; break; } ++c; }
                }
            } else {                                                    /* :400-410, Q14 */
                /* stable ascending rank under isless; p[k] = rank/normaliser; Distributions
                 * v0.20 DiscreteNonParametric sampler: cp=0,i=0; while cp<u && i<n: cp+=p[++i] */
                double normaliser = (double)(K * (K + 1) / 2);
                double cp = 0.0; int64_t i = 0;
                while (cp < u && i < K) {
                    int64_t k = i; ++i;
                    int64_t rank = 1;
                    for (int64_t j = 0; j < K; ++j)
                        rank += (jl_isless(rho[j], rho[k]) || (!jl_isless(rho[k], rho[j]) && j < k));
                    cp = cp + (double)rank / normaliser;
                }
                choice = (i > 0) ? i - 1 : 0;
            }
        }
        emit(w, t, ORC_EV_CHOICE, reinv, choice, injection);
        A_(w, reinv, choice) = injection; A_(w, reinv, K) = 0.0;        /* :358-359 */
        double fundval = orc_dotw(&H_(w, choice, 0), K, &P_(w, 0, t), 1, M);     /* :360 */
        double newcapital = fundval + injection;
        for (int64_t n = 0; n < N; ++n) {                               /* :363-366 */
            double sv = S_(w, choice, n) * fundval;
            if (n == reinv) sv = injection;
            S_(w, choice, n) = sv / newcapital;
        }
        double *bo = buy_values + (*nbuy) * M;
        for (int64_t m = 0; m < M; ++m) bo[m] = ((H_(w, choice, m) * P_(w, m, t)) / fundval) * injection; /* :367-368 */
        buy_funds[*nbuy] = choice; ++*nbuy;
    }
    free(rho); free(vt);
    return rc;
}

/* trackvolume!(volume, ::SellMarketOrder, t), functions.jl:555-558 */
static void trackvolume_sell(orc_world *w, int64_t t, const double *sale_values, int64_t norders)
{
    for (int64_t m = 0; m < w->M; ++m) {
        double net = 0.0;
        for (int64_t r = 0; r < norders; ++r) net = net + sale_values[r * w->M + m];
        w->volume[m + w->M * (t - 1)] = w->volume[m + w->M * (t - 1)] - net;
    }
}

/* modelrun, functions.jl:446-491 */
int orc_modelrun(orc_world *w, int64_t t_first, int64_t t_last, int selector, int flags)
{
    int64_t K = w->K, M = w->M, N = w->N;
    if (selector < 0 || selector > 3) return ORC_ERR_BAD_SELECTOR;       /* Q9 */
    w->faithful = (flags & ORC_FAITHFUL_HISTORY) != 0;
    if (!w->log_built) build_log(w);
    double *z = (double *)zalloc(8 * M);
    int64_t *reviewers = (int64_t *)zalloc(8 * N), *div_inv = (int64_t *)zalloc(8 * N), *div_fund = (int64_t *)zalloc(8 * N);
    double *sale = (double *)zalloc(8 * N * M), *cash = (double *)zalloc(8 * N);
    double *buy = (double *)zalloc(8 * N * M), *shares = (double *)zalloc(8 * N * M);
    int64_t *buy_funds = (int64_t *)zalloc(8 * N);
    int rc = ORC_OK;
    for (int64_t t = t_first; t <= t_last && rc == ORC_OK; ++t) {
        marketmove(w, t, draw_z_mkt(w, t));                               /* :458 */
        for (int64_t m = 0; m < M; ++m) z[m] = draw_z_stock(w, t, m);
        stockmove(w, t, z);                                               /* :459 */
        if (w->faithful) for (int64_t k = 0; k < K; ++k) orc_fundrevalue(w, k);  /* :460 */
        for (int64_t k = 0; k < K; ++k) {
            double v = w->faithful ? V_(w, k, t) : nav(w, k, t);
            w->nav_close[k] = v;
            if (w->trace_on) w->tr_nav_open[k + K * (t - 1)] = v;
        }
        /* drawreviewers :147-151 */
        int64_t nrev = 0;
        for (int64_t n = 0; n < N; ++n)
            if (draw_u_review(w, t, n) < w->p.reviewprobability) { reviewers[nrev++] = n; emit(w, t, ORC_EV_REVIEWER, n, 0, 0.0); }
        /* perfreview :157-176 */
        int64_t nd = 0;
        for (int64_t i = 0; i < nrev; ++i) {
            int64_t rev = reviewers[i], h = w->horizon[rev];
            for (int64_t f = 0; f < K; ++f) {
                if (!(A_(w, rev, f) > 0.0)) continue;
                double valchange = (w->nav_close[f] - nav(w, f, t - h)) / nav(w, f, h);
                if (valchange < w->threshold[rev]) { div_inv[nd] = rev; div_fund[nd] = f; ++nd; emit(w, t, ORC_EV_DIVEST, rev, f, 0.0); }
            }
        }
        if (nd > 0) {                                                     /* :464 */
            orc_liquidate(w, t, div_inv, div_fund, nd, sale);             /* :465 */
            trackvolume_sell(w, t, sale, nd);                             /* :467 */
            executeorder_sell_phase(w, t, sale, nd, cash);                /* :468 */
            for (int64_t r = 0; r < nd; ++r) emit(w, t, ORC_EV_CASHOUT, div_inv[r], div_fund[r], cash[r]);
            orc_disburse_cash(w, div_inv, div_fund, nd, cash);            /* :469 */
            if (w->trace_on) for (int64_t m = 0; m < M; ++m) w->tr_p_sell[m + M * (t - 1)] = P_(w, m, t);
            int any_dead = 0;                                             /* :471 */
            for (int64_t k = 0; k < K && !any_dead; ++k) {
                double s = 0.0;
                for (int64_t n = 0; n < N; ++n) s = s + S_(w, k, n);
                if (s == 0.0) any_dead = 1;
            }
            if (any_dead) {
                int64_t nbuy = 0;
                rc = orc_respawn(w, t, buy, buy_funds, &nbuy);            /* :472 */
                if (rc != ORC_OK) break;
                if (!(flags & ORC_NO_VOLUME_QUIRK)) trackvolume_sell(w, t, sale, nd);   /* :474, Q3 */
                executeorder_buy_phase(w, t, buy, nbuy, shares, 1);       /* :475 */
                disburse_shares_at(w, buy_funds, nbuy, shares, t);        /* :476 */
            }
            if (w->trace_on) for (int64_t m = 0; m < M; ++m) w->tr_p_resp[m + M * (t - 1)] = P_(w, m, t);
            if (w->faithful) for (int64_t k = 0; k < K; ++k) orc_fundrevalue(w, k);     /* :479 */
            for (int64_t k = 0; k < K; ++k) w->nav_close[k] = nav(w, k, t);
            int any_cash = 0;                                             /* :481 */
            for (int64_t n = 0; n < N && !any_cash; ++n) any_cash = (A_(w, n, K) > 0.0);
            if (any_cash) {
                int64_t nbuy = 0;
                rc = orc_reinvest(w, t, selector, buy, buy_funds, &nbuy); /* :482 */
                if (rc != ORC_OK) break;
                if (!(flags & ORC_NO_VOLUME_QUIRK)) trackvolume_sell(w, t, sale, nd);   /* :484, Q3 */
                executeorder_buy_phase(w, t, buy, nbuy, shares, 2);       /* :485 */
                disburse_shares_at(w, buy_funds, nbuy, shares, t);        /* :486 */
            }
        }
    }
    free(z); free(reviewers); free(div_inv); free(div_fund); free(sale); free(cash); free(buy); free(shares); free(buy_funds);
    w->err = rc;
    return rc;
}

/* funds.value as the reference leaves it after a run that ended at t_now:
 * every fund's row was last rewritten with its final holdings; column t_now is nav_close. */
void orc_finalise_fund_value(orc_world *w, int64_t t_now)
{
    for (int64_t k = 0; k < w->K; ++k) {
        orc_fundrevalue(w, k);
        V_(w, k, t_now) = w->nav_close[k];
    }
}

/* ------------------------------------------------------------------------------------------------
 * Func.calc_stylisedfacts, functions.jl:604-648 -- the array-valued outputs.
 * PARITY UNPINNED: the reference has no tests for these functions (TODOs at functions.jl:547,554,
 * 560,572) and the statistics come from packages whose sources are not under /root/reference:
 * StatsBase v0.30.0 (Manifest.toml:506-510) `pacf` (default method :regression = per lag l the
 * OLS fit of x[i] on [1, x[i-1..i-l]], i = l+1..n, solved as cholesky(X'X) \ X'y, last
 * coefficient), `kurtosis` (excess, population moments about the sample mean), `percentile`
 * (= Statistics.quantile(v, p/100): f0 = (n-1)p, i = trunc(f0)+1, h = f0-trunc(f0),
 * v[i] + h (v[i+1]-v[i]) on the sorted vector) and Statistics.cor (sums of centred products,
 * clamped to [-1,1]).  Restated from their published algorithms; all sums sequential, ascending.
 * Quirks kept: lossgainratio skips another perfwindow[end] returns of the already trimmed series
 * (functions.jl:575,578) and its cutoff is the 5th percentile of |x|.
 * ---------------------------------------------------------------------------------------------- */
static void sf_returns(const double *v, int64_t stride, int64_t T, int64_t W, double *x)
{   /* assetreturns, functions.jl:500-515: (v[i+1]-v[i])/v[i], i = W+1..T-1 (1-based); then demean :517-535 */
    const int64_t n = T - W - 1;
    double s = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        const double a = v[(W + j) * stride], b = v[(W + j + 1) * stride];
        x[j] = (b - a) / a;
        s = s + x[j];
    }
    const double mean = s / (double)n;
    for (int64_t j = 0; j < n; ++j) x[j] = x[j] - mean;
}

static double sf_pacf_lag(const double *x, int64_t n, int64_t l)
{   /* StatsBase pacf_regress!: regressors [1, x[i-1], ..., x[i-l]], response x[i], rows i = l+1..n */
    enum { MAXD = 65 };
    double G[MAXD][MAXD], rhs[MAXD];
    const int64_t d = l + 1;
    for (int64_t a = 0; a < d; ++a) {
        for (int64_t b = a; b < d; ++b) {
            double s = 0.0;
            for (int64_t i = l; i < n; ++i) s = s + (a ? x[i - a] : 1.0) * (b ? x[i - b] : 1.0);
            G[a][b] = G[b][a] = s;
        }
        double s = 0.0;
        for (int64_t i = l; i < n; ++i) s = s + (a ? x[i - a] : 1.0) * x[i];
        rhs[a] = s;
    }
    /* Cholesky G = L L', in place in the lower triangle */
    for (int64_t j = 0; j < d; ++j) {
        double s = G[j][j];
        for (int64_t k = 0; k < j; ++k) s = s - G[j][k] * G[j][k];
        const double ljj = sqrt(s);
        G[j][j] = ljj;
        for (int64_t i = j + 1; i < d; ++i) {
            double t = G[i][j];
            for (int64_t k = 0; k < j; ++k) t = t - G[i][k] * G[j][k];
            G[i][j] = t / ljj;
        }
    }
    for (int64_t i = 0; i < d; ++i) {          /* L y = rhs */
        double t = rhs[i];
        for (int64_t k = 0; k < i; ++k) t = t - G[i][k] * rhs[k];
        rhs[i] = t / G[i][i];
    }
    for (int64_t i = d - 1; i >= 0; --i) {     /* L' b = y */
        double t = rhs[i];
        for (int64_t k = i + 1; k < d; ++k) t = t - G[k][i] * rhs[k];
        rhs[i] = t / G[i][i];
    }
    return rhs[d - 1];
}

static int sf_cmp(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

int orc_stylisedfacts(const double *marketval, const double *stocksval, const double *stocksvolume,
                      int64_t M, int64_t T, int64_t W, int64_t maxlag, double pctile,
                      double *stock_pacf, double *market_pacf, double *stock_volacluster,
                      double *market_volacluster, double *kurtoses, double *lossgain, double *corrs)
{
    const int64_t n = T - W - 1;
    if (n < 2 || maxlag < 1 || maxlag > 64 || n <= 2 * maxlag + 1) return ORC_ERR_BAD_STATE;
    double *x = (double *)malloc(sizeof(double) * (size_t)n * 3);
    double *y = x + n, *a = y + n;
    for (int64_t s = 0; s <= M; ++s) {          /* s == M: the market index */
        if (s < M) sf_returns(stocksval + s, M, T, W, x); else sf_returns(marketval, 1, T, W, x);
        for (int64_t j = 0; j < n; ++j) y[j] = x[j] * x[j];   /* demeaned returns squared, :616-617 */
        for (int64_t l = 1; l <= maxlag; ++l) {               /* calc_pacf, :588-602 */
            const double c = sf_pacf_lag(x, n, l), v = sf_pacf_lag(y, n, l);
            if (s < M) { stock_pacf[s + M * (l - 1)] = c; stock_volacluster[s + M * (l - 1)] = v; }
            else       { market_pacf[l - 1] = c; market_volacluster[l - 1] = v; }
        }
        if (s == M) break;
        /* calc_kurtoses :538-545 -> StatsBase.kurtosis(v): moments about mean(v), minus 3 */
        {
            double sm = 0.0, c2 = 0.0, c4 = 0.0;
            for (int64_t j = 0; j < n; ++j) sm = sm + x[j];
            const double mu = sm / (double)n;
            for (int64_t j = 0; j < n; ++j) { const double z = x[j] - mu, z2 = z * z; c2 = c2 + z2; c4 = c4 + z2 * z2; }
            c2 = c2 / (double)n; c4 = c4 / (double)n;
            kurtoses[s] = c4 / (c2 * c2) - 3.0;
        }
        /* lossgainratio :572-586 on x[W+1:end] */
        {
            const int64_t n2 = n - W;
            if (n2 < 1) lossgain[s] = NAN;
            else {
                for (int64_t j = 0; j < n2; ++j) a[j] = fabs(x[W + j]);
                qsort(a, (size_t)n2, sizeof(double), sf_cmp);
                const double f0 = (double)(n2 - 1) * (pctile / 100.0), t0 = trunc(f0), h = f0 - t0;
                const int64_t i = (int64_t)t0;
                const double q = (h == 0.0) ? a[i] : a[i] + h * (a[i + 1] - a[i]);
                int64_t up = 0, down = 0;
                for (int64_t j = 0; j < n2; ++j) { up += x[W + j] > q; down += x[W + j] < -q; }
                lossgain[s] = (double)down / (double)up;
            }
        }
        /* volavolumecorr :560-570: cor(volume[s, W+2:end], |x|) */
        {
            const double *vol = stocksvolume + s + M * (W + 1);
            double sv = 0.0, sa = 0.0;
            for (int64_t j = 0; j < n; ++j) { sv = sv + vol[M * j]; sa = sa + fabs(x[j]); }
            const double mv = sv / (double)n, ma = sa / (double)n;
            double xx = 0.0, yy = 0.0, xy = 0.0;
            for (int64_t j = 0; j < n; ++j) {
                const double xi = vol[M * j] - mv, yi = fabs(x[j]) - ma;
                xx = xx + xi * xi; yy = yy + yi * yi; xy = xy + xi * yi;
            }
            const double mx = xx > yy ? xx : yy, mn = xx > yy ? yy : xx;
            double c = xy / mx / sqrt(mn / mx);
            if (c > 1.0) c = 1.0; else if (c < -1.0) c = -1.0;   /* clampcor; NaN passes through */
            corrs[s] = c;
        }
    }
    free(x);
    return ORC_OK;
}
