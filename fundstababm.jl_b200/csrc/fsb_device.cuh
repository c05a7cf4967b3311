// fsb_device.cuh -- device building blocks of the B200 engine: counter-based draws and the
// canonical reductions.  Compiled with -fmad=false: every fused multiply-add below is an
// explicit fma(); every other operation is a separately rounded IEEE-754 FP64 operation, so the
// results are bit-identical to the CPU oracle's on the same draws (DESIGN.md "Canonical order").
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fsb {

// ---- draw sites (counter = (block, a, site | b<<8, global replication id), key = seed) --------
enum : uint32_t {
    SITE_INIT_MKT = 1, SITE_INIT_BETA = 2, SITE_INIT_VOL = 3, SITE_INIT_STOCK = 4,
    SITE_INIT_IMPACT = 5, SITE_INIT_HORIZON = 6, SITE_INIT_THRESH = 7, SITE_INIT_CAPITAL = 8,
    SITE_INIT_FUNDPICK = 9, SITE_INIT_PSIZE = 10, SITE_INIT_PPICK = 11,
    SITE_MKT = 16, SITE_STOCK = 17, SITE_REVIEW = 18, SITE_SHUFFLE = 19, SITE_RPSIZE = 20,
    SITE_RPPICK = 21, SITE_CHOICE = 22
};

struct Rng {
    uint32_t k0, k1, rep;
};

// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ uint4 draw_block(const Rng &g, uint32_t site, uint32_t a, uint32_t b, uint32_t block)
{
    return philox4x32_10(make_uint4(block, a, site | (b << 8), g.rep), g.k0, g.k1);
}

// 52 random bits -> (0,1)
__device__ __forceinline__ double u52(uint32_t hi, uint32_t lo)
{
    const unsigned long long x = ((static_cast<unsigned long long>(hi) << 32) | lo) >> 12;
    return (static_cast<double>(x) + 0.5) * 0x1.0p-52;
}

// log on (0,1] from +,-,*,/,fma only (no libdevice transcendental), same steps as the oracle
__device__ __forceinline__ double det_log(double x)
{
    unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(x));
    long long e = static_cast<long long>((u >> 52) & 0x7ff) - 1023;
    u = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m = __longlong_as_double(static_cast<long long>(u));
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    const double f = m - 1.0;
    const double s = f / (m + 1.0);
    const double z = s * s;
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
    const double lm = 2.0 * s * p;
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double de = static_cast<double>(e);
    return fma(de, ln2_hi, fma(de, ln2_lo, lm));
}

// inverse normal CDF: Wichura AS241 (PPND16) on det_log
static __device__ __noinline__ double det_norminv(double u)
{
    const double q = u - 0.5;
    double r, num, den;
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
    r = sqrt(-det_log(r));
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
    const double val = num / den;
    return (q < 0.0) ? -val : val;
}

__device__ __forceinline__ double draw_uniform(const Rng &g, uint32_t site, uint32_t a, uint32_t b, uint32_t idx)
{
    const uint4 w = draw_block(g, site, a, b, idx >> 1);
    return (idx & 1) ? u52(w.z, w.w) : u52(w.x, w.y);
}
__device__ __forceinline__ double draw_normal(const Rng &g, uint32_t site, uint32_t a, uint32_t b, uint32_t idx)
{
    return det_norminv(draw_uniform(g, site, a, b, idx));
}
__device__ __forceinline__ uint32_t word_of(const uint4 &w, uint32_t i)
{
    return i == 0 ? w.x : (i == 1 ? w.y : (i == 2 ? w.z : w.w));
}
__device__ __forceinline__ uint32_t draw_discrete(const Rng &g, uint32_t site, uint32_t a, uint32_t b, uint32_t idx, uint32_t n)
{
    const uint4 w = draw_block(g, site, a, b, idx >> 2);
    return __umulhi(word_of(w, idx & 3), n);
}

// ---- canonical reductions: groups of 8 lanes ---------------------------------------------------
// A row of n doubles is read as double2 words; lane j (0..7) of a group owns words j, j+8, j+16, ...
// (elements 16i+2j, 16i+2j+1, ascending) as one chain from +0.0; the 8 chains are combined by the
// xor butterfly 4,2,1 (a+b == b+a bit-exactly, so every lane of the group ends with the same value).
// A warp holds 4 independent groups (4 rows at once, 128 contiguous bytes per group and load).
// All 32 lanes must call; `on` masks the loads of a group that has no row.
__device__ __forceinline__ double tree8(double a)
{
    a = a + __shfl_xor_sync(0xffffffffu, a, 4);
    a = a + __shfl_xor_sync(0xffffffffu, a, 2);
    a = a + __shfl_xor_sync(0xffffffffu, a, 1);
    return a;
}

// x, y: 16-byte aligned rows with an even number of allocated elements >= n
__device__ __forceinline__ double dot8(const double *x, const double *y, int n, int j, bool on = true)
{
    double acc = 0.0;
    if (on) {
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        const double2 *y2 = reinterpret_cast<const double2 *>(y);
        for (int e = 2 * j; e < n; e += 16) {
            const double2 xv = x2[e >> 1], yv = y2[e >> 1];
            acc = fma(xv.x, yv.x, acc);
            if (e + 1 < n) acc = fma(xv.y, yv.y, acc);
        }
    }
    return tree8(acc);
}
__device__ __forceinline__ double sum8(const double *x, int n, int j, bool on = true)
{
    double acc = 0.0;
    if (on) {
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        for (int e = 2 * j; e < n; e += 16) {
            const double2 xv = x2[e >> 1];
            acc = acc + xv.x;
            if (e + 1 < n) acc = acc + xv.y;
        }
    }
    return tree8(acc);
}
__device__ __forceinline__ double sumprod8(const double *x, const double *y, int n, int j, bool on = true)
{
    double acc = 0.0;
    if (on) {
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        const double2 *y2 = reinterpret_cast<const double2 *>(y);
        for (int e = 2 * j; e < n; e += 16) {
            const double2 xv = x2[e >> 1], yv = y2[e >> 1];
            acc = acc + xv.x * yv.x;
            if (e + 1 < n) acc = acc + xv.y * yv.y;
        }
    }
    return tree8(acc);
}

// ---- asynchronous-copy helpers (TMA unit, sm_90+) ----------------------------------------------
__device__ __forceinline__ void l2_prefetch_bulk(const void *gptr, unsigned bytes)  // bytes % 16 == 0
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// non-blocking probe (its ~150 cycles overlap with independent work when the result is used later)
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy completing on an mbarrier; dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one box of a 2-D tensor map (coordinates: c0 = element in the row, c1 = row) -> shared memory (128-byte aligned), on an mbarrier
__device__ __forceinline__ void tensor_g2s_2d(void *sdst, const void *tmap, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sdst)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Julia's isless on Float64: total order, -0.0 < +0.0, NaN last
__device__ __forceinline__ bool jl_isless(double a, double b)
{
    if (a != a) return false;
    if (b != b) return true;
    if (a < b) return true;
    if (a == b) return (__double_as_longlong(a) < 0) && !(__double_as_longlong(b) < 0);
    return false;
}

}  // namespace fsb
