// fsb_aux.cuh -- kernels around the step kernel: device-side Func.initialise, liquidation-log
// construction, ensemble statistics, full fund-value history for the drop-in download.
#pragma once
#include "fsb_device.cuh"
#include "fsb_state.h"

namespace fsb {

// ------------------------------------------------------------------------------------------------
// Func.initialise on the device (src/functions.jl:417-430; draw sites per SURVEY.md 3.2).
// One CTA per replication.  Dynamic shared memory: (W+2) doubles market returns + K doubles.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fsb_init_kernel(const DevState s)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *mret = reinterpret_cast<double *>(smem_raw);  // [W+2] market return entering column t
    double *capk = mret + (s.W + 2);                      // [K] fund capital
    const long long r = blockIdx.x;
    const long long g = s.rep_offset + r;
    const DevPoint &pt = s.points[s.reps_per_point > 0 ? static_cast<int>(g / s.reps_per_point) : 0];
    const double *tab = s.tables;
    Rng rng{static_cast<uint32_t>(s.seed), static_cast<uint32_t>(s.seed >> 32), static_cast<uint32_t>(g)};
    const int K = s.K, M = s.M, N = s.N, W = s.W, Mp = s.Mp, tid = threadIdx.x, NT = blockDim.x;
    double *H = s.H + r * s.h_stride;
    double *P = s.P + r * s.p_stride;
    double *beta = s.beta + r * Mp, *vol = s.vol + r * Mp, *impact = s.impact + r * Mp;
    double *market = s.market + r * s.mkt_len;
    int *pos = s.pos + r * N, *horizon = s.horizon + r * N;
    double *stake = s.stake + r * N, *amount = s.amount + r * N, *threshold = s.threshold + r * N;

    // marketinit! (:19-27)
    if (tid == 0) {
        double mk = pt.marketstartval;
        if (s.keep_history) market[0] = mk;
        for (int t = 2; t <= W + 1; ++t) {
            const double z = draw_normal(rng, SITE_INIT_MKT, 0, 0, static_cast<uint32_t>(t));
            const double g2 = (1.0 + pt.drift) + pt.marketvol * z;
            const double mn = mk * g2;
            mret[t] = (mn - mk) / mk;
            mk = mn;
            if (s.keep_history) market[t - 1] = mk;
        }
        if (!s.keep_history) market[mslot(s, W + 1)] = mk;
        s.mkt_peak[r] = mk;
        s.mkt_dd[r] = 0.0;
        s.err[r] = 0;
        s.log_n[r] = 0;
        for (int q = 0; q < kCnt; ++q) s.cnt[r * kCnt + q] = 0;
    }
    for (int q = tid; q < kFlowBins; q += NT) s.flow_hist[r * kFlowBins + q] = 0;
    // zero the whole price / volume / holdings storage of this replication
    for (long long q = tid; q < s.p_stride; q += NT) P[q] = 0.0;
    if (s.volume)
        for (long long q = tid; q < static_cast<long long>(s.T) * Mp; q += NT) s.volume[r * static_cast<long long>(s.T) * Mp + q] = 0.0;
    for (long long q = tid; q < s.h_stride; q += NT) H[q] = 0.0;
    __syncthreads();
    // betainit! (:48-52), stockvolinit! (:55-59), stockimpactinit! (:72-76), stockvalueinit! (:61-70)
    for (int m = tid; m < Mp; m += NT) {
        if (m >= M) { beta[m] = 0.0; vol[m] = 0.0; impact[m] = 0.0; continue; }
        const double b = draw_normal(rng, SITE_INIT_BETA, 0, 0, static_cast<uint32_t>(m)) * pt.betastd + pt.betamean;
        const double v = tab[pt.vol_off + draw_discrete(rng, SITE_INIT_VOL, 0, 0, static_cast<uint32_t>(m), static_cast<uint32_t>(pt.n_vol))];
        beta[m] = b;
        vol[m] = v;
        impact[m] = tab[pt.impact_off + draw_discrete(rng, SITE_INIT_IMPACT, 0, 0, static_cast<uint32_t>(m), static_cast<uint32_t>(pt.n_impact))];
        double prev = pt.stockstartval;
        P[m] = prev;
        for (int t = 2; t <= W + 1; ++t) {
            const double z = draw_normal(rng, SITE_INIT_STOCK, static_cast<uint32_t>(t), 0, static_cast<uint32_t>(m));
            const double a = (1.0 + mret[t] * b) * prev;
            const double c2 = (v * z) * prev;
            prev = a + c2;
            P[static_cast<long long>(t - 1) * Mp + m] = prev;
        }
    }
    // invhorizoninit! (:78-82), invthreshinit! (:84-88), invassetinit! (:91-102)
    for (int n = tid; n < N; n += NT) {
        horizon[n] = pt.perf_lo + static_cast<int>(draw_discrete(rng, SITE_INIT_HORIZON, 0, 0, static_cast<uint32_t>(n), static_cast<uint32_t>(pt.perf_hi - pt.perf_lo + 1)));
        threshold[n] = draw_normal(rng, SITE_INIT_THRESH, 0, 0, static_cast<uint32_t>(n)) * pt.thresholdstd + pt.thresholdmean;
        amount[n] = static_cast<double>(pt.cap_lo + static_cast<int>(draw_discrete(rng, SITE_INIT_CAPITAL, 0, 0, static_cast<uint32_t>(n), static_cast<uint32_t>(pt.cap_hi - pt.cap_lo + 1))));
        pos[n] = (n < K) ? n : static_cast<int>(draw_discrete(rng, SITE_INIT_FUNDPICK, 0, 0, static_cast<uint32_t>(n), static_cast<uint32_t>(K)));
    }
    __syncthreads();
    // fundcapitalinit! (:104-107): column sums over investors in ascending order
    for (int k = tid; k < K; k += NT) {
        double c2 = 0.0;
        for (int n = 0; n < N; ++n)
            if (pos[n] == k) c2 = c2 + amount[n];
        capk[k] = c2;
        s.cap0[r * K + k] = c2;
        s.nav_open[r * K + k] = 0.0;
        s.nav_close[r * K + k] = 0.0;
        s.live_row[r * K + k] = k;
    }
    __syncthreads();
    // fundstakeinit! (:111-117)
    for (int n = tid; n < N; n += NT) stake[n] = amount[n] / capk[pos[n]];
    __syncthreads();
    // stakeless[k] = (sum(stakes[k,:]) == 0), the respawn trigger of :471, kept incrementally by the step kernel
    for (int k = tid; k < K; k += NT) {
        double c2 = 0.0;
        for (int n = 0; n < N; ++n)
            if (pos[n] == k) c2 = c2 + stake[n];
        s.stakeless[r * K + k] = (c2 == 0.0) ? 1 : 0;
    }
    // fundholdinit! (:119-133): each thread owns a fund's row, picks applied in draw order
    for (int k = tid; k < K; k += NT) {
        const int nk = pt.psize_lo + static_cast<int>(draw_discrete(rng, SITE_INIT_PSIZE, 0, 0, static_cast<uint32_t>(k), static_cast<uint32_t>(pt.psize_hi - pt.psize_lo + 1)));
        double *row = H + static_cast<long long>(k) * Mp;
        bool nz = false;
        for (int j = 0; j < nk; ++j) {
            const int st = static_cast<int>(draw_discrete(rng, SITE_INIT_PPICK, static_cast<uint32_t>(k), 0, static_cast<uint32_t>(j), static_cast<uint32_t>(M)));
            const double hn = row[st] + (capk[k] / static_cast<double>(nk)) / P[st];
            row[st] = hn;
            nz = nz || (hn != 0.0);
        }
        s.hzero[r * K + k] = nz ? 0 : 1;
    }
}

// ------------------------------------------------------------------------------------------------
// liquidation log, src/functions.jl:448-456 (Q4: sized with K).  One CTA per replication,
// warp per fund.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fsb_build_log_kernel(const DevState s)
{
    const long long r = blockIdx.x;
    const int K = s.K, M = s.M, N = s.N, Mp = s.Mp;
    const int lane = threadIdx.x & 31, j = lane & 7, ngroups = blockDim.x >> 3, gid = threadIdx.x >> 3;
    const double *H = s.H + r * s.h_stride;
    const double *P1 = s.P + r * s.p_stride;  // column 1
    const double *impact = s.impact + r * Mp;
    const int *pos = s.pos + r * N;
    const double *stake = s.stake + r * N;
    const long long lb = r * s.cap_log;
    for (int kb = (threadIdx.x >> 5) * 4; kb < K; kb += ngroups) {  // warp-uniform: 4 funds per warp
        const int k = kb + (lane >> 3);
        const bool on = k < K;
        const double *row = H + static_cast<long long>(on ? k : kb) * Mp;
        // calc_liquidity (:324-330)
        const double tot = sumprod8(row, P1, M, j, on);
        double acc = 0.0;
        int nst = 0;
        if (on)
            for (int e = 2 * j; e < M; e += 16) {
                acc = acc + ((row[e] * P1[e]) / tot) * impact[e];
                nst += (row[e] > 0.0);
                if (e + 1 < M) { acc = acc + ((row[e + 1] * P1[e + 1]) / tot) * impact[e + 1]; nst += (row[e + 1] > 0.0); }
            }
        const double liq = tree8(acc);
        int ninv = 0;
        if (on)
            for (int n = j; n < N; n += 8) ninv += (pos[n] == k && stake[n] > 0.0);
#pragma unroll
        for (int sft = 4; sft >= 1; sft >>= 1) {
            nst += __shfl_xor_sync(0xffffffffu, nst, sft);
            ninv += __shfl_xor_sync(0xffffffffu, ninv, sft);
        }
        if (on && j == 0) {
            s.log_liquidated[lb + k] = 0;
            s.log_replacedby[lb + k] = 0;
            s.log_size[lb + k] = s.cap0[r * K + k];
            s.log_liquidity[lb + k] = liq;
            s.log_failtime[lb + k] = 0;
            s.log_ninv[lb + k] = ninv;
            s.log_nstocks[lb + k] = nst;
            s.live_row[r * K + k] = k;
        }
    }
    (void)gid;
    if (threadIdx.x == 0) s.log_n[r] = K;
}

// ------------------------------------------------------------------------------------------------
// ensemble statistics: integer counters and histograms per sweep point (order independent)
// ------------------------------------------------------------------------------------------------
__global__ void fsb_stats_kernel(const DevState s, unsigned long long *out, long long words_per_point)
{
    const long long r = blockIdx.x;
    const long long g = s.rep_offset + r;
    const int pt = s.reps_per_point > 0 ? static_cast<int>(g / s.reps_per_point) : 0;
    unsigned long long *o = out + static_cast<long long>(pt) * words_per_point;
    unsigned long long *h_fail = o + FSB_STATS_HEADER;
    unsigned long long *h_liq = h_fail + (s.T + 1);
    unsigned long long *h_dd = h_liq + FSB_STATS_LIQ_BINS;
    unsigned long long *h_flow = h_dd + FSB_STATS_DD_BINS;
    const int rows = s.log_n[r];
    const long long lb = r * s.cap_log;
    int nliq = 0;
    for (int q = threadIdx.x; q < rows; q += blockDim.x) {
        if (s.log_liquidated[lb + q]) {
            ++nliq;
            const int ft = s.log_failtime[lb + q];
            if (ft >= 0 && ft <= s.T) atomicAdd(&h_fail[ft], 1ULL);
        }
    }
    for (int q = threadIdx.x; q < kFlowBins; q += blockDim.x) {
        const int v = s.flow_hist[r * kFlowBins + q];
        if (v) atomicAdd(&h_flow[q], static_cast<unsigned long long>(v));
    }
    __shared__ int nliq_sh, nan_sh;
    if (threadIdx.x == 0) { nliq_sh = 0; nan_sh = 0; }
    __syncthreads();
    if (nliq) atomicAdd(&nliq_sh, nliq);
    for (int k = threadIdx.x; k < s.K; k += blockDim.x) {
        const double v = s.nav_close[r * s.K + k];
        if (v != v) nan_sh = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int *cnt = s.cnt + r * kCnt;
        atomicAdd(&o[FSB_ST_REPS], 1ULL);
        atomicAdd(&o[FSB_ST_LIQUIDATIONS], static_cast<unsigned long long>(nliq_sh));
        atomicAdd(&o[FSB_ST_FUNDS_EVER], static_cast<unsigned long long>(rows));
        atomicAdd(&o[FSB_ST_FLOOR_HITS], static_cast<unsigned long long>(cnt[CNT_FLOOR]));
        atomicAdd(&o[FSB_ST_DIVESTMENTS], static_cast<unsigned long long>(cnt[CNT_DIVEST]));
        atomicAdd(&o[FSB_ST_REVIEWS], static_cast<unsigned long long>(cnt[CNT_REVIEWS]));
        atomicAdd(&o[FSB_ST_REINVESTMENTS], static_cast<unsigned long long>(cnt[CNT_REINVEST]));
        atomicAdd(&o[FSB_ST_STEPS], static_cast<unsigned long long>(cnt[CNT_STEPS]));
        if (s.err[r] & ~FSB_REP_TRACE_OVERFLOW) atomicAdd(&o[FSB_ST_ERR_REPS], 1ULL);
        if (nan_sh) atomicAdd(&o[FSB_ST_NAN_REPS], 1ULL);
        atomicAdd(&h_liq[min(nliq_sh, FSB_STATS_LIQ_BINS - 1)], 1ULL);
        const double dd = s.mkt_dd[r];
        int bin = static_cast<int>(dd * FSB_STATS_DD_BINS);
        bin = max(0, min(FSB_STATS_DD_BINS - 1, bin));
        atomicAdd(&h_dd[bin], 1ULL);
    }
}

// ------------------------------------------------------------------------------------------------
// funds.value for the drop-in download (src/functions.jl:266-271 leaves V[k,:] = holdings[k,:]'*P
// for every column; column t_now holds nav_close).  Needs keep_history.  out: [K][T] row-major.
// ------------------------------------------------------------------------------------------------
__global__ void fsb_fund_value_kernel(const DevState s, long long r, int t_now, double *out)
{
    const int j = threadIdx.x & 7;
    const long long wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = (gridDim.x * blockDim.x) >> 5;
    const double *H = s.H + r * s.h_stride;
    const double *P = s.P + r * s.p_stride;
    const long long total = static_cast<long long>(s.K) * s.T;
    for (long long qb = wglobal * 4; qb < total; qb += nw * 4) {  // warp-uniform: 4 (fund, column) pairs per warp
        const long long q = qb + ((threadIdx.x & 31) >> 3);
        const bool on = q < total;
        const int k = on ? static_cast<int>(q / s.T) : 0, col = on ? static_cast<int>(q % s.T) : 0;  // col 0-based
        double v = dot8(H + static_cast<long long>(k) * s.Mp, P + static_cast<long long>(col) * s.Mp, s.M, j, on);
        if (col == t_now - 1) v = s.nav_close[r * s.K + k];
        if (on && j == 0) out[q] = v;
    }
}

// ---- debug kernels for cross-checks against the oracle ------------------------------------------
__global__ void fsb_debug_draws_kernel(unsigned long long seed, uint32_t rep, uint32_t site, uint32_t a, uint32_t b,
                                       uint32_t idx0, int n, int kind, uint32_t nd, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rng rng{static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), rep};
    const uint32_t idx = idx0 + static_cast<uint32_t>(i);
    double v;
    if (kind == 0) v = draw_normal(rng, site, a, b, idx);
    else if (kind == 1) v = draw_uniform(rng, site, a, b, idx);
    else v = static_cast<double>(draw_discrete(rng, site, a, b, idx, nd));
    out[i] = v;
}

__global__ void fsb_debug_dot_kernel(const double *x, const double *y, int n, double *out)
{
    const double v = dot8(x, y, n, threadIdx.x & 7);
    if (threadIdx.x == 0) *out = v;
}

}  // namespace fsb
