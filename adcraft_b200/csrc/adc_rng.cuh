// Counter-based draws of the B200 AdCraft step: Philox4x32-10 + deterministic samplers.
//
// The "tape function" (DESIGN.md): every random quantity of (env, keyword, step) is a pure
// function of a Philox counter, so results do not depend on grid shape or GPU count.
//   key = (seed_lo, seed_hi);  ctr = (index, step, stream<<28 | agent<<20 | keyword, env)
// All floating-point arithmetic uses explicit round-to-nearest intrinsics (no implicit FMA
// contraction, no fast-math), so a CPU that evaluates the same expressions with IEEE fmaf gets
// bit-identical values -- that is what lets the tests compare integer outcomes exactly.
//
// Reference behaviour sampled here:
//   competitor bid   around(max(|Laplace(loc,scale)|,0),2)   adcraft/synthetic_kw_helpers.py:104-113
//   revenue          around(max(N(mu,sd),0.01),2)            adcraft/synthetic_kw_helpers.py:66-70
//   volume           round(max(N(mu,sd),0)) half away        src/lib.rs:314-325
//   click / conv     U <= p                                   adcraft/synthetic_kw_helpers.py:73-77
//   impressions      Binomial(n, thresholded sigmoid)         src/lib.rs:69-76,92-105
//   explicit cost    clamp(sqrt(b)/4 + 2.2 + N(0,1e-10+sqrt(b)/6), 0, 4.4)   src/lib.rs:53-67
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace adc {

enum : uint32_t { ST_AUCTION = 0u, ST_UNIT = 1u, ST_REVENUE = 2u, ST_PHANTOM = 3u, ST_IDEAL = 4u, ST_COST = 5u, ST_BIDDERS = 6u };

struct PhiloxKey {
    uint32_t k0, k1;
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// Philox with the first round split off: c1 (step), c2 (stream|kw) and c3 (env) are fixed for a
// whole (env, keyword, step) unit, so half of round 1 is computed once per unit (philox_pre) and
// the per-draw work is 1 + 2*9 multiplies (philox_from_pre).  Same function as philox4x32_10.
struct PhiloxPre {
    uint32_t n0, n1, x3;  // n0 = hi(M1*c2)^c1^k0, n1 = lo(M1*c2), x3 = c3^k1
};

__device__ __forceinline__ PhiloxPre philox_pre(uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                uint32_t k1)
{
    PhiloxPre p;
    p.n0 = __umulhi(0xCD9E8D57u, c2) ^ c1 ^ k0;
    p.n1 = 0xCD9E8D57u * c2;
    p.x3 = c3 ^ k1;
    return p;
}

// 32x32 -> 64 multiply as ONE IMAD.WIDE.U32 (the plain C++ forms sometimes compile to IMAD.HI + IMAD)
__device__ __forceinline__ void mulwide(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
    unsigned long long p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p));
}

__device__ __forceinline__ uint4 philox_from_pre(uint32_t idx, uint32_t n0, uint32_t n1, uint32_t x3,
                                                 uint32_t k0, uint32_t k1)
{
    uint32_t c0 = n0, c1 = n1, c2, c3;
    mulwide(0xD2511F53u, idx, c2, c3);
    c2 ^= x3;
#pragma unroll
    for (int r = 1; r < 10; ++r) {
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
        uint32_t h0, l0, h1, l1;
        mulwide(0xD2511F53u, c0, h0, l0);
        mulwide(0xCD9E8D57u, c2, h1, l1);
        c0 = h1 ^ c1 ^ k0;
        c2 = h0 ^ c3 ^ k1;
        c1 = l1;
        c3 = l0;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ uint32_t stream_word(uint32_t stream, uint32_t agent, uint32_t kw)
{
    return (stream << 28) | (agent << 20) | kw;
}

constexpr float kLn2f = 0.693147182f;

// -ln((w31 + 0.5) / 2^31): Exp(1) variate from a 31-bit uniform integer.
// a = 2*w31+1 is normalised (a << clz) to x = an/2^32 in [0.5,1); -ln(x) comes from a 128-entry
// chord table on the 7 bits below the leading one (error <= 7.6e-6, tools/gen_neglog_table.py):
//   e = clz * ln2 + T[idx] - S[idx] * lo,   lo = low 24 bits.
// 12 instructions instead of the ~26 of a range-reduced polynomial log.  `tab` may point at the
// global copy below or at a shared-memory copy (the hot kernel stages it).
__device__ const float2 kNeglogTab[128] = {
#include "adc_neglog_table.inc"
};

__device__ __forceinline__ float neglog_norm(uint32_t a, const float2 *tab)
{
    uint32_t lz;  // a is odd, so bfind.shiftamt (= clz for a != 0) is the normalising shift
    asm("bfind.shiftamt.u32 %0, %1;" : "=r"(lz) : "r"(a));
    const uint32_t an = a << lz;
    const float2 ts = tab[(an >> 24) & 0x7Fu];
    const float lo = __uint2float_rn(an & 0x00FFFFFFu);
    const float inner = __fmaf_rn(-lo, ts.y, ts.x);
    return __fmaf_rn(__uint2float_rn(lz), kLn2f, inner);
}

__device__ __forceinline__ float neglog_u31(uint32_t w31, const float2 *tab)
{
    return neglog_norm(2u * w31 + 1u, tab);
}

// Standard normal from one word: sign bit + 31-bit two-sided tail probability t = (2*w31+1)/2^32
// = P(|Z| > z).  Like the Exp(1) sampler: a = 2*w31+1 is normalised and z = sqrt(2)*erfcinv(t) is
// read off a minimax chord table on (clz, the 7 bits below the leading one), error <= 1.6e-6
// (tools/gen_znorm_table.py): 12 instructions instead of the ~60 of log + erfinv polynomials.
// The 32 KB table lives in global memory; the rows in use (clz 0..3 take 94 % of the draws) stay
// in L1.
__device__ const float2 kZnormTab[32 * 128] = {
#include "adc_znorm_table.inc"
};

__device__ __forceinline__ float znorm(uint32_t w)
{
    uint32_t a, lz;  // a = 2 * (w & 0x7FFFFFFF) + 1 (mod 2^32)
    asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(a) : "r"(w));
    asm("bfind.shiftamt.u32 %0, %1;" : "=r"(lz) : "r"(a));
    const uint32_t an = a << lz;
    // the leading one of `an` is folded into the table base: (an >> 24) = 128 + idx
    const float2 ts = __ldg(kZnormTab + (int)(lz * 128u + (an >> 24)) - 128);
    const float lo = __uint2float_rn(an & 0x00FFFFFFu);
    const float z = __fmaf_rn(-lo, ts.y, ts.x);
    return __uint_as_float(__float_as_uint(z) ^ (w & 0x80000000u));
}

// exp(x) in float64, explicit fma only (explicit keywords' thresholded sigmoid, once per unit).
__device__ __forceinline__ double exp_det(double x)
{
    if (x != x) return x;
    if (x > 709.0) return __longlong_as_double(0x7FF0000000000000LL);
    if (x < -700.0) return 0.0;
    const double kf = rint(__dmul_rn(x, 1.4426950408889634));
    double r = __fma_rn(-kf, 6.93147180369123816490e-01, x);
    r = __fma_rn(-kf, 1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0;
    p = __fma_rn(p, r, 1.0 / 479001600.0);
    p = __fma_rn(p, r, 1.0 / 39916800.0);
    p = __fma_rn(p, r, 1.0 / 3628800.0);
    p = __fma_rn(p, r, 1.0 / 362880.0);
    p = __fma_rn(p, r, 1.0 / 40320.0);
    p = __fma_rn(p, r, 1.0 / 5040.0);
    p = __fma_rn(p, r, 1.0 / 720.0);
    p = __fma_rn(p, r, 1.0 / 120.0);
    p = __fma_rn(p, r, 1.0 / 24.0);
    p = __fma_rn(p, r, 1.0 / 6.0);
    p = __fma_rn(p, r, 0.5);
    p = __fma_rn(p, r, 1.0);
    p = __fma_rn(p, r, 1.0);
    const long long k = (long long)kf;
    const double s = __longlong_as_double((k + 1023LL) << 52);
    return __dmul_rn(p, s);
}

// 15 instructions: 2 * (w0 & 0x7FFFFFFF) + 1 (mod 2^32) as one multiply-add, the sign of the
// Laplace branch XORed into the scale's sign bit.
__device__ __forceinline__ int laplace_cents(uint32_t w0, float loc, float scale,
                                             const float2 *tab = kNeglogTab)
{
    uint32_t a;
    asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(a) : "r"(w0));
    const float e = neglog_norm(a, tab);
    const float s = __uint_as_float(__float_as_uint(scale) ^ (w0 & 0x80000000u));
    const float x = __fmaf_rn(s, e, loc);
    return __float2int_rn(__fmul_rn(fabsf(x), 100.0f));
}

// One competitor's bid of the default ImplicitKeyword: signed, un-rounded Laplace(loc, scale)
// (synthetic_kw_classes.py:670-688).
__device__ __forceinline__ double laplace_signed(uint32_t w0, float loc, float scale, const float2 *tab = kNeglogTab)
{
    uint32_t a;
    asm("mad.lo.u32 %0, %1, 2, 1;" : "=r"(a) : "r"(w0));
    const float e = neglog_norm(a, tab);
    const float s = __uint_as_float(__float_as_uint(scale) ^ (w0 & 0x80000000u));
    return (double)__fmaf_rn(s, e, loc);
}

__device__ __forceinline__ int revenue_cents(uint32_t w, float mean, float sd)
{
    const float v = __fmaf_rn(sd, znorm(w), mean);
    const int c = __float2int_rn(__fmul_rn(v, 100.0f));
    return c < 1 ? 1 : c;
}

__device__ __forceinline__ long long volume_draw(uint32_t w, double mean, double sd)
{
    double v = __dadd_rn(mean, __dmul_rn(sd, (double)znorm(w)));
    if (!(v > 0.0)) v = 0.0;
    return (long long)round(v);
}

__device__ __forceinline__ double clampd(double x, double lo, double hi)
{
    return x < lo ? lo : (x > hi ? hi : x);
}

__device__ __forceinline__ double threshold_sigmoid(double bid, double thresh_in, double intercept,
                                                    double slope)
{
    const double halver = 2.0 + 1e-10;
    const double thresh = __ddiv_rn(clampd(__dmul_rn(halver, thresh_in), 0.0, 1.0), halver);
    const double ex = exp_det(__dmul_rn(-slope, __dsub_rn(bid, intercept)));
    const double r = __ddiv_rn(1.0, __dadd_rn(1.0, ex));
    const double a = __dadd_rn(1.0, __dmul_rn(2.0, thresh));
    return clampd(__dsub_rn(__dmul_rn(a, r), thresh), 0.0, 1.0);
}

// P(w * 2^-32 <= p)  <=>  w <= floor(p * 2^32), saturated.
__device__ __forceinline__ uint32_t prob_threshold(double p)
{
    if (!(p > 0.0)) return 0u;
    const double t = floor(__dmul_rn(p, 4294967296.0));
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)t;
}

// Implicit keywords take click and conversion from ONE word cc (two auctions share a Philox call):
// click <=> cc <= T1 = prob_threshold(ctr); given a click cc is uniform on [0, T1], so
// u_conv = (cc + 0.5) / (T1 + 1) is a fresh uniform and conversion <=> u_conv <= cvr (the
// reference's comparison, synthetic_kw_helpers.py:77) <=> cc < T2, T2 = #{v <= T1 : u_conv(v) <= cvr}.
__device__ __forceinline__ unsigned long long conv_threshold(uint32_t t1, double cvr)
{
    const double n = __dadd_rn((double)t1, 1.0);
    double est = ceil(__dsub_rn(__dmul_rn(cvr, n), 0.5));
    if (!(est > 0.0)) est = 0.0;
    if (est > n) est = n;
    long long v = (long long)est;
    while (v > 0 && __ddiv_rn(__dadd_rn((double)(v - 1), 0.5), n) > cvr) --v;
    while ((double)v < n && __ddiv_rn(__dadd_rn((double)v, 0.5), n) <= cvr) ++v;
    return (unsigned long long)v;
}

__device__ __forceinline__ double explicit_cost(uint32_t w3, double bid)
{
    const double xs = __dsqrt_rn(bid);
    const double sd = __dadd_rn(1e-10, __ddiv_rn(xs, 6.0));
    const double mean = __dadd_rn(__ddiv_rn(xs, 4.0), 4.4 / 2.0);
    const double c = __dadd_rn(mean, __dmul_rn(sd, (double)znorm(w3)));
    return clampd(c, 0.0, 4.4);
}

// round(np.maximum(bid, 0.01), 2) in integer cents (adcraft/gymnasium_kw_env.py:215).
__device__ __forceinline__ int bid_to_cents(double bid)
{
    double b = bid > 0.01 ? bid : 0.01;
    if (!(b == b)) b = 0.01;
    double c = rint(__dmul_rn(b, 100.0));
    if (c > 2.0e9) c = 2.0e9;
    return (int)c;
}

// ------------------------------------------------------------------------------------------
// Free-running implicit keywords: the O(clicks) tape function.
//
// The reference draws a competitor bid for every auction (helpers:104-113) although only the
// clicked auctions' prices are ever used.  Here ONE 32-bit uniform R_j per auction decides the
// nested events  win <=> R_j < T1,  click <=> R_j < T2 = T1 ctr,  conversion <=> R_j < T3 = T2 cvr,
// with T1 = 2^32 P(round(|Laplace|, 2) < win_cents), and each clicked auction draws its price from
// the competitor-bid law conditioned on that event (cost_cents2).  R_j is bit-sliced: auctions
// come in groups of 32 (bit p of a word = auction 32 g + p); level l (0 = MSB of R) of group g is
// word (l & 3) of Philox call 8 g + (l >> 2) of the AUCTION stream, so a thread compares 32
// auctions with the three thresholds by a few word operations per level and stops at the first
// level where every auction is decided (about log2(3 * 32) + 1.3 levels).
// ------------------------------------------------------------------------------------------
struct Unit2 {
    uint32_t t1, t2, t3;  // thresholds mod 2^32 ...
    uint32_t full;        // ... bit i set: threshold i+1 is 2^32 ("always")
    uint32_t h1;          // half-width (2^-31 units) of window 1, <= 2^31
    uint32_t a1, a2;      // e^-hi of the two windows, 2^-32 units
    float L, b;           // |loc|, scale
    int W;                // cents the competitor must stay below
};

__device__ __forceinline__ unsigned long long rint_u64(double x)
{
    return x <= 0.0 ? 0ull : (unsigned long long)rint(x);
}

__device__ __forceinline__ Unit2 unit2_make(double loc, double scale, double ctr, double cvr, int win_cents)
{
    const double ymax = __ddiv_rn(__dsub_rn((double)win_cents, 0.5), 100.0);
    const double L = fabs(loc);
    const double b = scale > 1e-9 ? scale : 1e-9;
    const double d = __dsub_rn(L, ymax);
    const double lo1 = __ddiv_rn(d > 0.0 ? d : 0.0, b), hi1 = __ddiv_rn(__dadd_rn(L, ymax), b);
    const double e_lo1 = exp_det(-lo1), e_hi1 = exp_det(-hi1);
    double l1 = __dsub_rn(e_lo1, e_hi1), l2 = 0.0, e_hi2 = 1.0;
    if (!(l1 > 0.0)) l1 = 0.0;
    if (ymax > L) {
        e_hi2 = exp_det(-__ddiv_rn(__dsub_rn(ymax, L), b));
        l2 = __dsub_rn(1.0, e_hi2);
    }
    unsigned long long h1 = rint_u64(__dmul_rn(l1, 2147483648.0));
    const unsigned long long h2 = rint_u64(__dmul_rn(l2, 2147483648.0));
    unsigned long long T1 = h1 + h2;
    if (T1 > 4294967296ull) T1 = 4294967296ull;
    if (h1 > T1) h1 = T1;
    unsigned long long T2 = rint_u64(__dmul_rn((double)T1, clampd(ctr, 0.0, 1.0)));
    if (T2 > T1) T2 = T1;
    unsigned long long T3 = rint_u64(__dmul_rn((double)T2, clampd(cvr, 0.0, 1.0)));
    if (T3 > T2) T3 = T2;
    const unsigned long long a1 = rint_u64(__dmul_rn(e_hi1, 4294967296.0));
    const unsigned long long a2 = rint_u64(__dmul_rn(e_hi2, 4294967296.0));
    Unit2 u;
    u.t1 = (uint32_t)T1; u.t2 = (uint32_t)T2; u.t3 = (uint32_t)T3;
    u.full = (T1 >> 32 ? 1u : 0u) | (T2 >> 32 ? 2u : 0u) | (T3 >> 32 ? 4u : 0u);
    u.h1 = (uint32_t)h1;  // <= 2^31
    u.a1 = a1 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)a1;
    u.a2 = a2 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)a2;
    u.L = (float)L; u.b = (float)b; u.W = win_cents;
    return u;
}

// Price of a clicked auction in cents.  t = e^-E is uniform on (e^-hi1, e^-lo1) [x = L - b E] united
// with (e^-hi2, 1) [x = L + b E]; the word picks a point of that union (all integer up to -ln).
__device__ __forceinline__ int cost_cents2(uint32_t w, uint32_t t1, bool t1_full, uint32_t h1, uint32_t a1, uint32_t a2,
                                           float L, float b, int W, const float2 *tab)
{
    const uint32_t s = t1_full ? w : __umulhi(w, t1);
    const bool left = s < h1;
    const uint32_t off = left ? s : s - h1;
    // base + 2 off + 1 saturated at 2^32 - 1 (off < 2^31): a wrapped 32-bit sum is below its base
    const uint32_t base = left ? a1 : a2;
    uint32_t t = base + 2u * off + 1u;
    if (t < base) t = 0xFFFFFFFFu;
    const float e = neglog_norm(t | 1u, tab);
    const float x = __fmaf_rn(left ? -b : b, e, L);
    int c = __float2int_rn(__fmul_rn(fabsf(x), 100.0f));
    c = min(c, W - 1);
    return max(c, 0);
}

// One group of up to 32 auctions against the three thresholds: `active` = bits that are auctions.
// Returns the win / click / conversion masks.  Lazy: reads levels only while some auction is still
// tied with some threshold's prefix.
struct Masks3 {
    uint32_t win, click, conv;
};

// all ones when bit `pos` (a compile-time position) of x is set: one BFE.S32 of a one-bit field
template <int dummy = 0>
__device__ __forceinline__ uint32_t sign_bit_mask(uint32_t x, int pos)
{
    int m;
    asm("bfe.s32 %0, %1, %2, 1;" : "=r"(m) : "r"((int)x), "r"(pos));
    return (uint32_t)m;
}

__device__ __forceinline__ Masks3 group_masks(uint32_t active, uint32_t g, uint32_t t1, uint32_t t2, uint32_t t3,
                                              uint32_t full, uint32_t n0, uint32_t n1, uint32_t x3, uint32_t k0,
                                              uint32_t k1)
{
    // E_i: auctions whose R matches T_i on the levels read so far; L_i: auctions known to be below T_i
    uint32_t e1 = (full & 1u) ? 0u : active, e2 = (full & 2u) ? 0u : active, e3 = (full & 4u) ? 0u : active;
    uint32_t l1 = (full & 1u) ? active : 0u, l2 = (full & 2u) ? active : 0u, l3 = (full & 4u) ? active : 0u;
#pragma unroll 1
    for (uint32_t q = 0; q < 8u && (e1 | e2 | e3) != 0u; ++q) {
        const uint4 w = philox_from_pre(8u * g + q, n0, n1, x3, k0, k1);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t r = ws[i];
            // m_i = all ones when the threshold's bit at this level is 1 (a sign-extending one-bit field
            // extract at a fixed position: the thresholds move up four levels per call):
            //   bit 1: auctions with r = 0 drop below (L |= E & ~r), those with r = 1 stay tied (E &= r)
            //   bit 0: auctions with r = 1 rise above (E &= ~r)
            const uint32_t m1 = sign_bit_mask(t1, 31 - i), m2 = sign_bit_mask(t2, 31 - i), m3 = sign_bit_mask(t3, 31 - i);
            l1 |= e1 & ~r & m1; e1 &= ~(r ^ m1);
            l2 |= e2 & ~r & m2; e2 &= ~(r ^ m2);
            l3 |= e3 & ~r & m3; e3 &= ~(r ^ m3);
        }
        t1 <<= 4; t2 <<= 4; t3 <<= 4;
    }
    Masks3 m;
    m.win = l1; m.click = l2; m.conv = l3;
    return m;
}

// numpy >= 2 keeps a float32 bid float32 through round(np.maximum(bid, 0.01), 2) (env:215), and
// searchsorted then compares float32(cents / 100) upcast to float64 with the float64 competitor bids
// (helpers:166-170): when the float32 value lies above the cent value the bid wins ties.  Returns 1
// for such a bid: the win test is `cents + bonus > competitor`.
__device__ __forceinline__ int bid_tie_bonus(int cents)
{
    const float f = __fdiv_rn((float)cents, 100.0f);
    return (double)f > __ddiv_rn((double)cents, 100.0) ? 1 : 0;
}

}  // namespace adc
