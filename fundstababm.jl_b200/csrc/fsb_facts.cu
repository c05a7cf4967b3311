// fsb_facts.cu -- Func.calc_stylisedfacts (src/functions.jl:604-648) on the device, from the price,
// volume and market histories a keep_history engine holds in HBM: per stock the partial
// autocorrelations of the demeaned returns and of their squares (lags 1..10, calc_pacf :588-602), the
// excess kurtosis (calc_kurtoses :538-545), the loss/gain ratio (lossgainratio :572-586) and the
// volume-|return| correlation (volavolumecorr :560-570); for the market index the two PACFs.
// Only M-vectors leave the device instead of the M x T histories (2 x 2.7 MB per replication).
//
// One thread per series (stock m, or the market index as series M), 128 series per CTA: thread m
// walks its series through time, so a warp reads 32 consecutive stocks of one price column -- the
// histories are [t][asset], asset fastest, and every load is coalesced.  Two launches, one per role:
// <0> the series' own statistics in three walks over the prices and one over the volume -- (1) the mean
// return, (2) lagged products of x, moments, the volume sums and the order statistics the percentile
// needs, (3) up/down counts against the percentile; <1> volatility clustering -- the lagged
// products of x^2 (the mean return comes from <0>).
//
// PACF by regression as StatsBase v0.30.0 does it (per lag l: OLS of x[i] on [1, x[i-1..i-l]],
// i = l+1..n, through a factorisation of X'X) -- but the (l+1)^2 sums of every lag are derived
// from 11 full-range lagged sums F_d = sum_s x[s] x[s-d] minus the few head / tail products each
// window leaves out, instead of 10 passes with up to 66 running sums.
#include <cuda_runtime.h>

#include <algorithm>

#include "fsb_state.h"

namespace fsb {

constexpr int kFactLags = 10;   // lags carried in registers (the reference's maxlag default, :605)
constexpr int kFactSel = 128;   // order statistics a thread can keep for the percentile
constexpr int kFactThreads = 128;  // series per CTA; a replication's M + 1 series spread over gridDim.y CTAs

struct FactOut {  // device arrays in the layouts of include/fsb.h, indexed from the first replication of the launch
    double *pacf, *volc;    // [n_reps][maxlag][M]
    double *mpacf, *mvolc;  // [n_reps][maxlag]
    double *kurt, *lg, *corr;  // [n_reps][M]
    double *mean;              // [n_reps][M + 1] scratch: the mean returns, launch <0> -> launch <1>
};

struct FactSeries {  // where a thread's series lives
    const double *p;       // first value (column 1)
    long long stride;      // doubles between consecutive days
    const double *vol;     // volume, same stride (null for the market index)
};

// Window corrections of the lagged sums, built once per series: the regression of lag l runs over rows
// i = l..n-1 (0-based), so sum_i z[i-a] z[i-b] = F[d] (d = b-a, all s = d..n-1) minus the head products
// s = d..l-a-1 and the tail products s = n-a..n-1.  hd[d][m] = the first m head products of lag d,
// td[d][a] = the last a tail products of lag d, hs[m] / ts[a] the same for the plain sums.
struct FactEdges {
    double hd[kFactLags + 1][kFactLags + 1], td[kFactLags + 1][kFactLags + 1], hs[kFactLags + 1], ts[kFactLags + 1];
};
__device__ __noinline__ void fact_edges(FactEdges &e, const double *head, const double *tail /* tail[k] = z[n-1-k] */)
{
    e.hs[0] = e.ts[0] = 0.0;
    for (int m = 0; m < kFactLags; ++m) { e.hs[m + 1] = e.hs[m] + head[m]; e.ts[m + 1] = e.ts[m] + tail[m]; }
    for (int d = 0; d <= kFactLags; ++d) {
        e.hd[d][0] = e.td[d][0] = 0.0;
        for (int m = 0; m + d < kFactLags; ++m) {
            e.hd[d][m + 1] = e.hd[d][m] + head[d + m] * head[m];
            e.td[d][m + 1] = e.td[d][m] + tail[m] * tail[m + d];
        }
    }
}

// last coefficient of the regression of z[i] on [1, z[i-1], ..., z[i-l]], rows i = l..n-1, from the full-range
// sums F[d] = sum_{s=d..n-1} z[s] z[s-d], S = sum z and the window corrections
__device__ __noinline__ double fact_pacf_lag(const double *F, double S, const FactEdges &e, int n, int l)
{
    double G[kFactLags + 1][kFactLags + 1], rhs[kFactLags + 1];
    for (int a = 0; a <= l; ++a) {
        const double sa = S - e.hs[l - a] - e.ts[a];  // sum_{i=l..n-1} z[i-a]
        for (int b = a; b <= l; ++b) {
            const int d = b - a;
            const double c = F[d] - e.hd[d][l - b] - e.td[d][a];
            if (a == 0) { if (b > 0) rhs[b] = c; }
            else G[b][a] = c;  // lower triangle, b >= a >= 1
        }
        if (a == 0) rhs[0] = sa; else G[a][0] = sa;
    }
    G[0][0] = static_cast<double>(n - l);
    // X'X = L D L' (unit lower triangular L, no square roots, one reciprocal per column instead of a division per
    // entry); L y = rhs; the last unknown of L' b = D^-1 y is y_last / D_last
    const int dm = l + 1;
    double v[kFactLags + 1], dlast = 1.0;
    for (int j = 0; j < dm; ++j) {
        double dj = G[j][j];
        for (int k = 0; k < j; ++k) {
            v[k] = G[j][k] * G[k][k];  // L_jk D_k
            dj -= G[j][k] * v[k];
        }
        G[j][j] = dj;
        dlast = dj;
        const double inv = 1.0 / dj;
        for (int i = j + 1; i < dm; ++i) {
            double t = G[i][j];
            for (int k = 0; k < j; ++k) t -= G[i][k] * v[k];
            G[i][j] = t * inv;
        }
    }
    for (int i = 1; i < dm; ++i) {
        double t = rhs[i];
        for (int k = 0; k < i; ++k) t -= G[i][k] * rhs[k];
        rhs[i] = t;
    }
    return rhs[dm - 1] / dlast;
}

// Keeps the qcap smallest values seen so far in a max-heap (the percentile needs its top two).
__device__ __forceinline__ void fact_heap_insert(double *sel, int &nsel, int qcap, double a)
{
    if (nsel < qcap) {  // sift up
        int k = nsel++;
        while (k > 0) {
            const int up = (k - 1) >> 1;
            const double pv = sel[up];
            if (!(pv < a)) break;
            sel[k] = pv;
            k = up;
        }
        sel[k] = a;
    } else {  // replaces the largest: sift down
        int k = 0;
        for (;;) {
            int c = 2 * k + 1;
            if (c >= nsel) break;
            double cv = sel[c];
            if (c + 1 < nsel) {
                const double c2 = sel[c + 1];
                if (c2 > cv) { cv = c2; ++c; }
            }
            if (!(cv > a)) break;
            sel[k] = cv;
            k = c;
        }
        sel[k] = a;
    }
}

// Empties a lane's candidate buffer into its heap; a candidate may have been overtaken since it was staged.
constexpr int kFactStage = 8;
__device__ __noinline__ void fact_heap_flush(double *sel, int &nsel, int qcap, const double *stg, int &nstg)
{
    for (int i = 0; i < nstg; ++i) {
        const double a = stg[i];
        if (nsel < qcap || a < sel[0]) fact_heap_insert(sel, nsel, qcap, a);
    }
    nstg = 0;
}

// Walks a series through time: fn(j, return j, volume j) for j = 0..n-1.  A thread's loads are one per day and
// each would wait a full HBM round trip if issued where it is used; the loads of kFactAhead days are issued
// together, ahead of the arithmetic on them.  (An additional prefetch.global.L2 of the lines 24 days ahead
// measured -5 %: the remaining stalls are not HBM latency.)
template <bool kVol, int kFactAhead, class Fn>
__device__ __forceinline__ void fact_walk(const double *__restrict__ p0, const double *__restrict__ v0, long long stride, int n, Fn &&fn)
{
    double pa = p0[0];
    int j0 = 0;
    for (; j0 + kFactAhead <= n; j0 += kFactAhead) {  // whole groups: no guards, so the unrolled window shifts are renames
        double pb[kFactAhead], vv[kFactAhead];
#pragma unroll
        for (int u = 0; u < kFactAhead; ++u) {
            pb[u] = p0[static_cast<long long>(j0 + u + 1) * stride];
            vv[u] = kVol ? v0[static_cast<long long>(j0 + u) * stride] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kFactAhead; ++u) {
            fn(j0 + u, (pb[u] - pa) / pa, vv[u]);
            pa = pb[u];
        }
    }
#pragma unroll 1
    for (; j0 < n; ++j0) {
        const double pb = p0[static_cast<long long>(j0 + 1) * stride];
        fn(j0, (pb - pa) / pa, kVol ? v0[static_cast<long long>(j0) * stride] : 0.0);
        pa = pb;
    }
}

// kRole 0: the series' own statistics (PACF of x, kurtosis, loss/gain ratio, volume correlation);
// kRole 1: volatility clustering (PACF of y = x^2).  Two launches: each role carries 11 lagged sums and a 10-deep
// window in registers instead of both (168 registers, 12 warps per SM), so 4 CTAs fit per SM.
template <int kRole>
__global__ void __launch_bounds__(kFactThreads, 4)
fsb_facts_kernel(const DevState s, const long long rep_first, const int maxlag, const double pct, const FactOut out)
{
    const long long r = rep_first + blockIdx.x;
    const int M = s.M, T = s.T, W = s.W, Mp = s.Mp;
    const int n = T - W - 1;      // returns per series (assetreturns :500-515)
    const int n2 = n - W;         // returns the loss/gain ratio looks at (:575,578)
    const int nser = M + 1;
    const long long rb = blockIdx.x;
    // percentile(v, pct) = quantile(v, pct/100): f0 = (n2-1) p, i = trunc(f0), h = f0 - i (Statistics._quantile)
    int qi = 0, qcap = 0;
    double qh = 0.0;
    if (n2 >= 1) {
        const double f0 = static_cast<double>(n2 - 1) * (pct / 100.0), t0 = trunc(f0);
        qh = f0 - t0;
        qi = static_cast<int>(t0);
        qcap = min(qi + 2, n2);
    }
    for (int ser = blockIdx.y * kFactThreads + threadIdx.x; ser < nser; ser += gridDim.y * kFactThreads) {
        FactSeries fs;
        if (ser < M) {
            fs.p = s.P + r * s.p_stride + ser;
            fs.stride = Mp;
            fs.vol = s.volume + r * static_cast<long long>(T) * Mp + ser;
        } else {
            fs.p = s.market + r * s.mkt_len;
            fs.stride = 1;
            fs.vol = nullptr;
        }
        const double *p0 = fs.p + static_cast<long long>(W) * fs.stride;            // column W+1
        // column W+2 of the volume; the market index has none and reads its own values instead (sums unused), so that
        // one instantiation of the walk serves every series
        const double *v0 = fs.vol ? fs.vol + static_cast<long long>(W + 1) * fs.stride : p0;
        const int ostride = ser < M ? M : 1;
        // pass 1: the mean return (launch <0> walks for it and leaves it for launch <1>, which follows on the same stream)
        double mean;
        if constexpr (kRole == 0) {
            double sr = 0.0;
            fact_walk<false, 8>(p0, v0, fs.stride, n, [&](int, double ret, double) { sr += ret; });
            mean = sr / static_cast<double>(n);
            out.mean[rb * nser + ser] = mean;
        } else {
            mean = out.mean[rb * nser + ser];
        }
        // pass 2: the 11 full-range lagged sums of z (x or y) with a window of the last 10 z, and the first 10 z
        double F[kFactLags + 1], zw[kFactLags], head[kFactLags];
#pragma unroll
        for (int d = 0; d <= kFactLags; ++d) F[d] = 0.0;
#pragma unroll
        for (int d = 0; d < kFactLags; ++d) zw[d] = 0.0;
        for (int d = 0; d < kFactLags; ++d) head[d] = 0.0;
        double sz = 0.0;
        auto lagged = [&](int j, double z) {
            sz += z;
            F[0] += z * z;
#pragma unroll
            for (int d = 1; d <= kFactLags; ++d) F[d] += z * zw[d - 1];
#pragma unroll
            for (int d = kFactLags - 1; d > 0; --d) zw[d] = zw[d - 1];
            zw[0] = z;
            if (j < kFactLags) head[j] = z;
        };
        auto solve = [&](double *o) {  // the sums leave their registers for the per-lag solves (local memory from here on)
            double Fl[kFactLags + 1], tail[kFactLags];
#pragma unroll
            for (int d = 0; d <= kFactLags; ++d) Fl[d] = F[d];
#pragma unroll
            for (int d = 0; d < kFactLags; ++d) tail[d] = zw[d];
            FactEdges ed;
            fact_edges(ed, head, tail);
            for (int l = 1; l <= maxlag; ++l) o[static_cast<long long>(l - 1) * ostride] = fact_pacf_lag(Fl, sz, ed, n, l);
        };
        if constexpr (kRole == 1) {
            fact_walk<false, 8>(p0, v0, fs.stride, n, [&](int j, double ret, double) {
                const double x = ret - mean;
                lagged(j, x * x);
            });
            solve(ser < M ? out.volc + rb * maxlag * M + ser : out.mvolc + rb * maxlag);
        } else {
            // ... plus the fourth moment, sum |x|, the raw volume sums and the order statistics of |x| the percentile needs
            double sel[kFactSel], stg[kFactStage];
            int nstg = 0;
            double s4 = 0.0, sabs = 0.0, sv = 0.0, svv = 0.0, sva = 0.0;
            int nsel = 0;
            const bool stock = fs.vol != nullptr;
            fact_walk<true, 4>(p0, v0, fs.stride, n, [&](int j, double ret, double v) {
                const double x = ret - mean, a = fabs(x), y = x * x;
                s4 += y * y;
                sabs += a;
                sv += v;
                svv += v * v;
                sva += v * a;
                lagged(j, x);
                // Candidates wait in a small per-lane buffer and all lanes of the warp empty theirs together, when the first
                // one is full: inserted one by one the 32 lanes qualify on different days, and the warp would run the sift
                // for one or two of them nearly every day.
                if (stock && j >= W && qcap > 0) {
                    if (nsel < qcap || a < sel[0]) stg[nstg++] = a;
                    if (__any_sync(__activemask(), nstg == kFactStage)) fact_heap_flush(sel, nsel, qcap, stg, nstg);
                }
            });
            fact_heap_flush(sel, nsel, qcap, stg, nstg);
            solve(ser < M ? out.pacf + rb * maxlag * M + ser : out.mpacf + rb * maxlag);
            if (!stock) continue;
            const double nn = static_cast<double>(n);
            // kurtosis (StatsBase: population moments about the sample mean, minus 3).  The series is demeaned already;
            // its residual mean (~1e-19) moves the moments by less than their last bit and is dropped.
            {
                const double c2 = F[0] / nn, c4 = s4 / nn;
                out.kurt[rb * M + ser] = c4 / (c2 * c2) - 3.0;
            }
            // pass 3: loss/gain counts
            double q = 0.0;
            if (n2 >= 1) {  // sorted a[qi], a[qi+1] = the two largest of the qi+2 smallest (the heap's root and its larger child)
                double hi = sel[0], lo = sel[0];
                if (qcap == qi + 2) {
                    lo = sel[1];
                    if (qcap > 2 && sel[2] > lo) lo = sel[2];
                }
                q = (qh == 0.0) ? lo : lo + qh * (hi - lo);
            }
            int up = 0, down = 0;
            fact_walk<false, 8>(p0, v0, fs.stride, n, [&](int j, double ret, double) {
                const double x = ret - mean;
                if (j >= W) { up += x > q; down += x < -q; }
            });
            out.lg[rb * M + ser] =
                n2 >= 1 ? static_cast<double>(down) / static_cast<double>(up) : __longlong_as_double(0x7ff8000000000000LL);
            // cor(volume, |x|) from the raw sums of pass 2: sum (v - mv)^2 = svv - sv^2 / n, sum (|x| - ma)^2 = F_0 - sabs^2 / n,
            // sum (v - mv)(|x| - ma) = sva - sv sabs / n.  None of the three cancels badly: volumes are sparse spikes
            // (mean^2 << mean square), |x| has mean^2 / mean square ~ 2 / pi, and the error of the cross term is
            // ~1e-16 (mv / sd v)(ma / sd |x|) of the normaliser.  A constant volume series (the reference's NaN: no trade
            // ever touched the stock) must not turn into rounding noise: below n eps of its raw sum the centred sum is 0.
            double xx = svv - sv * sv / nn;
            if (!(xx > svv * nn * 2.3e-16)) xx = (xx != xx) ? xx : 0.0;
            const double yy = F[0] - sabs * sabs / nn, xy = sva - sv * sabs / nn;
            const double mx = xx > yy ? xx : yy, mn = xx > yy ? yy : xx;
            double c = xy / mx / sqrt(mn / mx);
            if (c > 1.0) c = 1.0; else if (c < -1.0) c = -1.0;
            out.corr[rb * M + ser] = c;
        }
    }
}

// launched by fsb_stylised_facts (fsb_abi.cu); out: seven device arrays carved from one allocation of
// facts_doubles_per_rep(M, maxlag) doubles per replication, in the order of FactOut
int facts_max_lag() { return kFactLags; }
int facts_max_sel() { return kFactSel; }
long long facts_doubles_per_rep(int M, int maxlag) { return 2LL * maxlag * M + 2LL * maxlag + 3LL * M + (M + 1); }
cudaError_t facts_launch(const DevState &d, cudaStream_t st, long long rep_first, int n_reps, int maxlag, double pct, double *base)
{
    const long long M = d.M, nr = n_reps;
    FactOut o;
    o.pacf = base;
    o.volc = o.pacf + nr * maxlag * M;
    o.mpacf = o.volc + nr * maxlag * M;
    o.mvolc = o.mpacf + nr * maxlag;
    o.kurt = o.mvolc + nr * maxlag;
    o.lg = o.kurt + nr * M;
    o.corr = o.lg + nr * M;
    o.mean = o.corr + nr * M;
    const dim3 grid(static_cast<unsigned>(n_reps), static_cast<unsigned>(std::min((d.M + 1 + kFactThreads - 1) / kFactThreads, 64)));
    fsb_facts_kernel<0><<<grid, kFactThreads, 0, st>>>(d, rep_first, maxlag, pct, o);
    fsb_facts_kernel<1><<<grid, kFactThreads, 0, st>>>(d, rep_first, maxlag, pct, o);
    return cudaGetLastError();
}

}  // namespace fsb
