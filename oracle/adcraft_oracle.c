/* TEST INFRASTRUCTURE -- CPU oracle for the BiddingSimulation.step hot path.
 * See adcraft_oracle.h for scope and parity status (PINNED against the reference
 * Python through tests/golden/ and tests/test_oracle_vs_reference.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off; contraction must stay off so
 * that the float32/float64 sampler arithmetic below is bit-identical to the CUDA
 * implementation, which is compiled with --fmad=false and uses explicit fma).
 *
 * Layout of this file
 *   1. Philox4x32-10 + the deterministic samplers ("tape function")
 *   2. restatement of the src/lib.rs helpers on the path
 *   3. the env step: volume split, lane loop, shared budget, early break
 *   4. drift, batched free-running driver (OpenMP) used as the CPU baseline
 */
#include "adcraft_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* 1. Philox4x32-10 (Salmon et al., SC'11; Random123 KATs in tests)           */
/* ------------------------------------------------------------------------- */
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

/* counter layout: c0 = index, c1 = step, c2 = stream<<28 | agent<<20 | kw, c3 = env */
enum { ST_AUCTION = 0, ST_UNIT = 1, ST_REVENUE = 2, ST_PHANTOM = 3, ST_IDEAL = 4, ST_COST = 5, ST_BIDDERS = 6 };

static void draw4(uint64_t seed, uint32_t env, uint32_t step, uint32_t agent, uint32_t kw,
                  uint32_t stream, uint32_t idx, uint32_t out[4])
{
    uint32_t ctr[4] = { idx, step, (stream << 28) | (agent << 20) | kw, env };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    orc_philox4x32_10(ctr, key, out);
}

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

#define ORC_LN2F 0.693147182f

/* -ln((w31 + 0.5) / 2^31) for a 31-bit uniform integer: an Exp(1) variate.
 * a = 2*w31+1 is normalised (a << clz) to x = an/2^32 in [0.5,1); -ln(x) is read off a 128-entry
 * chord table on the 7 bits below the leading one (error <= 7.6e-6, tools/gen_neglog_table.py):
 *   e = clz * ln2 + T[idx] - S[idx] * lo,   lo = low 24 bits. */
static const float ORC_NEGLOG_TAB[128][2] = {
#include "neglog_table.inc"
};

float orc_neglog_u31(uint32_t w31)
{
    const uint32_t a = 2u * w31 + 1u;
    const int lz = __builtin_clz(a);
    const uint32_t an = a << lz;
    const uint32_t idx = (an >> 24) & 0x7Fu;
    const uint32_t lo = an & 0x00FFFFFFu;
    const float inner = fmaf(-(float)lo, ORC_NEGLOG_TAB[idx][1], ORC_NEGLOG_TAB[idx][0]);
    return fmaf((float)lz, ORC_LN2F, inner);
}

/* Standard normal from one 32-bit word: sign bit + 31-bit two-sided tail probability
 * t = (2*w31+1)/2^32 = P(|Z| > z).  a = 2*w31+1 is normalised (a << clz) and z = sqrt(2)*erfcinv(t)
 * is read off a chord table on (clz, the 7 bits below the leading one), minimax-shifted
 * (error <= 1.6e-6, tools/gen_znorm_table.py):   z = T[clz][idx] - S[clz][idx] * lo,  lo = low 24 bits. */
static const float ORC_ZNORM_TAB[32 * 128][2] = {
#include "znorm_table.inc"
};

float orc_znorm(uint32_t w)
{
    const uint32_t a = 2u * w + 1u; /* = 2 * (w & 0x7FFFFFFF) + 1 (mod 2^32) */
    const int lz = __builtin_clz(a);
    const uint32_t an = a << lz;
    const uint32_t idx = (uint32_t)lz * 128u + ((an >> 24) & 0x7Fu);
    const uint32_t lo = an & 0x00FFFFFFu;
    const float z = fmaf(-(float)lo, ORC_ZNORM_TAB[idx][1], ORC_ZNORM_TAB[idx][0]);
    return (w >> 31) ? -z : z;
}

/* exp(x) in float64 with explicit fma (used once per unit by the explicit keyword's
 * thresholded sigmoid; libm exp is not bit-reproducible across CPU/GPU). */
double orc_exp(double x)
{
    if (x != x) return x;
    if (x > 709.0) return INFINITY;
    if (x < -700.0) return 0.0;
    double kf = rint(x * 1.4426950408889634);
    double r = fma(-kf, 6.93147180369123816490e-01, x);
    r = fma(-kf, 1.90821492927058770002e-10, r);
    double p = 1.0 / 6227020800.0; /* 1/13! */
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int64_t k = (int64_t)kf;
    uint64_t bits = (uint64_t)(k + 1023) << 52; /* 2^k, k in [-1010, 1023] */
    double s; memcpy(&s, &bits, 8);
    return p * s;
}

/* Competitor bid of the single-competitor ImplicitKeyword, in cents:
 * around(max(|Laplace(loc,scale)|, 0), 2)  (synthetic_kw_helpers.py:104-113). */
int32_t orc_laplace_cents(uint32_t w0, float loc, float scale)
{
    float e = orc_neglog_u31(w0 & 0x7FFFFFFFu);
    float s = (w0 >> 31) ? -scale : scale;
    float x = fmaf(s, e, loc);
    return (int32_t)lrintf(fabsf(x) * 100.0f);
}

/* around(max(N(mean,std), 0.01), 2) in cents (synthetic_kw_helpers.py:66-70). */
int32_t orc_revenue_cents(uint32_t w, float mean, float std)
{
    float v = fmaf(std, orc_znorm(w), mean);
    int32_t c = (int32_t)lrintf(v * 100.0f);
    return c < 1 ? 1 : c;
}

/* round_half_away(max(N(mean,std), 0))  (src/lib.rs:314-325). */
int64_t orc_volume(uint32_t w, double mean, double std)
{
    double v = mean + std * (double)orc_znorm(w);
    if (!(v > 0.0)) v = 0.0;
    return (int64_t)round(v);
}

/* ------------------------------------------------------------------------- */
/* 2. src/lib.rs helpers on the path                                           */
/* ------------------------------------------------------------------------- */
static inline double clampd(double x, double lo, double hi)
{   /* num::clamp: if x < lo {lo} else if x > hi {hi} else x */
    return x < lo ? lo : (x > hi ? hi : x);
}

/* src/lib.rs:92-105 + :290-300, exp replaced by orc_exp (<= 2 ulp from libm). */
double orc_threshold_sigmoid(double bid, double thresh_in, double intercept, double slope)
{
    double halver = 2.0 + 1e-10;
    double thresh = clampd(halver * thresh_in, 0.0, 1.0) / halver;
    double r = 1.0 / (1.0 + orc_exp(-slope * (bid - intercept)));
    return clampd((1.0 + 2.0 * thresh) * r - thresh, 0.0, 1.0);
}

/* P(u <= p) for u = w * 2^-32: w <= floor(p * 2^32), saturated to 2^32-1. */
uint32_t orc_prob_threshold(double p)
{
    if (!(p > 0.0)) return 0u; /* u <= 0 only for w == 0 */
    double t = floor(p * 4294967296.0);
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)t;
}

/* Implicit keywords draw click and conversion from ONE 32-bit word cc (two auctions share a
 * Philox call): click <=> cc <= T1 = orc_prob_threshold(ctr).  Given a click, cc is uniform on
 * [0, T1], so u_conv = (cc + 0.5) / (T1 + 1) is a uniform on (0,1) independent of everything the
 * reference has used so far; conversion <=> u_conv <= cvr (the reference's own comparison,
 * synthetic_kw_helpers.py:77) <=> cc < T2 with T2 = #{v in [0,T1] : fl((v+0.5)/(T1+1)) <= cvr}. */
double orc_conv_uniform(uint32_t cc, uint32_t t1)
{
    return ((double)cc + 0.5) / ((double)t1 + 1.0);
}

uint64_t orc_conv_threshold(uint32_t t1, double cvr)
{
    const double n = (double)t1 + 1.0;
    double est = ceil(cvr * n - 0.5);
    if (!(est > 0.0)) est = 0.0;
    if (est > n) est = n;
    int64_t v = (int64_t)est;
    while (v > 0 && ((double)(v - 1) + 0.5) / n > cvr) --v;
    while ((double)v < n && ((double)v + 0.5) / n <= cvr) ++v;
    return (uint64_t)v;
}

/* src/lib.rs:53-67 cost_create (constant 4.4) for one impression. */
double orc_explicit_cost(uint32_t w3, double bid)
{
    double xs = sqrt(bid);
    double sd = 1e-10 + xs / 6.0;
    double c = (xs / 4.0 + 4.4 / 2.0) + sd * (double)orc_znorm(w3);
    return clampd(c, 0.0, 4.4);
}

/* ndarray::sum on a contiguous slice = numeric_util::unrolled_fold (src/lib.rs:107-111) */
double orc_sum_array(const double *x, int64_t n)
{
    double p[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    int64_t i = 0;
    for (; i + 8 <= n; i += 8)
        for (int j = 0; j < 8; ++j) p[j] += x[i + j];
    double s = 0.0;
    s += p[0] + p[4];
    s += p[1] + p[5];
    s += p[2] + p[6];
    s += p[3] + p[7];
    for (; i < n; ++i) s += x[i];
    return s;
}

int32_t orc_bid_to_cents(double bid)
{   /* round(np.maximum(bid, 0.01), 2): rint(x*100)/100  (gymnasium_kw_env.py:215) */
    double b = bid > 0.01 ? bid : 0.01;
    if (!(b == b)) b = 0.01;
    double c = rint(b * 100.0);
    if (c > 2.0e9) c = 2.0e9;
    return (int32_t)c;
}

/* ------------------------------------------------------------------------- */
/* 2b. free-running implicit keywords: the O(clicks) tape function */
/* ------------------------------------------------------------------------- */
/* The reference draws one competitor bid per auction, c = around(|Laplace(loc,scale)|, 2)
 * (helpers:104-113), wins iff bid > c (helpers:166-177), then flips click and conversion coins
 * (helpers:73-77).  Only the clicked auctions' prices ever matter, so the free-running mode draws
 * exactly that: per auction j ONE 32-bit uniform R_j decides the nested events
 *     win  <=> R_j < T1 = P(c < win_cents) * 2^32
 *     click<=> R_j < T2 = T1 * ctr           conversion <=> R_j < T3 = T2 * cvr
 * and every CLICKED auction draws its price from the competitor-bid distribution conditioned on
 * c < win_cents (inverse CDF of the folded Laplace restricted to the winning window).  The joint law
 * of (impressions, clicks, conversions, costs) is the reference's; lost and unclicked auctions never
 * materialise a competitor bid.
 *
 * R_j is stored bit-sliced so that the GPU can evaluate 32 auctions per word operation: auctions
 * come in groups of 32 (g = j / 32, bit p = j % 32); level l (0 = most significant) of group g is
 * word (l & 3) of Philox call 8 g + (l >> 2) of the AUCTION stream; bit p of that word is bit
 * (31 - l) of R_j.  The GPU stops reading levels as soon as every auction of the group is decided;
 * this oracle assembles the full 32-bit R_j and compares. */
typedef struct {
    uint64_t T1, T2, T3; /* thresholds in [0, 2^32] */
    uint64_t h1;         /* half-width (2^-31 units) of window 1: x in (-ymax, min(L, ymax)) */
    uint32_t A1, A2;     /* e^-hi of the two windows, 2^-32 units */
    float L, b;          /* |loc|, scale */
    int32_t W;           /* cents the competitor must stay below (bid cents, +1 under the f32 tie rule) */
} orc_unit2;

static uint64_t rint_u64(double x) { return x <= 0.0 ? 0u : (uint64_t)rint(x); }

void orc_unit2_make(double loc, double scale, double ctr, double cvr, int32_t win_cents, orc_unit2 *u)
{
    const double ymax = ((double)win_cents - 0.5) / 100.0; /* round(|x| * 100) < W  <=>  |x| < (W - 0.5) / 100 */
    const double L = fabs(loc);
    const double b = scale > 1e-9 ? scale : 1e-9;
    const double d = L - ymax;
    const double lo1 = (d > 0.0 ? d : 0.0) / b, hi1 = (L + ymax) / b;
    const double e_lo1 = orc_exp(-lo1), e_hi1 = orc_exp(-hi1);
    double l1 = e_lo1 - e_hi1, l2 = 0.0, e_hi2 = 1.0;
    if (!(l1 > 0.0)) l1 = 0.0;
    if (ymax > L) { e_hi2 = orc_exp(-((ymax - L) / b)); l2 = 1.0 - e_hi2; }
    u->h1 = rint_u64(l1 * 2147483648.0);
    const uint64_t h2 = rint_u64(l2 * 2147483648.0);
    u->T1 = u->h1 + h2;
    if (u->T1 > 4294967296ull) u->T1 = 4294967296ull;
    if (u->h1 > u->T1) u->h1 = u->T1;
    const double c1 = clampd(ctr, 0.0, 1.0), c2 = clampd(cvr, 0.0, 1.0);
    u->T2 = rint_u64((double)u->T1 * c1);
    if (u->T2 > u->T1) u->T2 = u->T1;
    u->T3 = rint_u64((double)u->T2 * c2);
    if (u->T3 > u->T2) u->T3 = u->T2;
    const uint64_t a1 = rint_u64(e_hi1 * 4294967296.0), a2 = rint_u64(e_hi2 * 4294967296.0);
    u->A1 = a1 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)a1;
    u->A2 = a2 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)a2;
    u->L = (float)L; u->b = (float)b; u->W = win_cents;
}

/* -ln(a / 2^32) for an odd 32-bit a (the table sampler of orc_neglog_u31 on a given fixed-point value) */
static float neglog_fixed(uint32_t a)
{
    const int lz = __builtin_clz(a);
    const uint32_t an = a << lz;
    const uint32_t idx = (an >> 24) & 0x7Fu;
    const uint32_t lo = an & 0x00FFFFFFu;
    const float inner = fmaf(-(float)lo, ORC_NEGLOG_TAB[idx][1], ORC_NEGLOG_TAB[idx][0]);
    return fmaf((float)lz, ORC_LN2F, inner);
}

/* Price of a clicked auction in cents: the competitor bid given that it lost to `W`.
 * t = e^-E is uniform on the union of (e^-hi1, e^-lo1) [x = L - b E, left of the mode or the
 * mirrored tail] and (e^-hi2, 1) [x = L + b E, right of the mode]; w picks a point of that union. */
int32_t orc_cost_cents(uint32_t w, const orc_unit2 *u)
{
    const uint64_t s = u->T1 >= 4294967296ull ? (uint64_t)w : (((uint64_t)w * u->T1) >> 32);
    const int left = s < u->h1;
    const uint64_t off = left ? s : s - u->h1;
    uint64_t t = (uint64_t)(left ? u->A1 : u->A2) + 2u * off + 1u;
    if (t > 0xFFFFFFFFull) t = 0xFFFFFFFFull;
    const float e = neglog_fixed((uint32_t)t | 1u);
    const float x = fmaf(left ? -u->b : u->b, e, u->L);
    int32_t c = (int32_t)lrintf(fabsf(x) * 100.0f);
    if (c > u->W - 1) c = u->W - 1;
    return c < 0 ? 0 : c;
}

/* the 32 uniforms R_j of group g (auctions 32 g .. 32 g + 31) */
static void group_uniforms(uint64_t seed, uint32_t env, uint32_t step, uint32_t agent, uint32_t kw,
                           uint32_t g, uint32_t R[32])
{
    uint32_t w[4];
    memset(R, 0, 32 * sizeof(uint32_t));
    for (uint32_t q = 0; q < 8; ++q) {
        draw4(seed, env, step, agent, kw, ST_AUCTION, 8u * g + q, w);
        for (uint32_t i = 0; i < 4; ++i) {
            const uint32_t l = 4u * q + i;
            for (uint32_t p = 0; p < 32; ++p) R[p] |= ((w[i] >> p) & 1u) << (31u - l);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* 3. one env step                                                             */
/* ------------------------------------------------------------------------- */
typedef struct {
    int mode; /* 0 tape, 1 philox */
    const orc_tape *tape;
    uint64_t seed; uint32_t env, step, agent;
    orc_record *rec;
    const int32_t *floor_cents; /* shared auctions: highest rival bid per keyword, or NULL */
    const int32_t *win_cents;   /* free-running: cents the competitor must stay below, or NULL = bid_cents */
} draw_src;

typedef struct { /* per-keyword running cursors for one env step */
    int64_t auction; /* auctions evaluated so far (day-level ordinal) */
    int64_t n_click, n_conv, n_rev, n_cost;
    int64_t n_clk; /* free-running implicit: clicked auctions so far, accepted or not (price-draw rank) */
} kw_cursor;

#define MAX_LANE_STACK 4096

static int lane_run(const orc_keywords *kw, int k, int t, int32_t bid_cents, double *remaining,
                    int alias, int64_t n, draw_src *src, kw_cursor *cur, orc_result *out,
                    uint32_t thr_click, uint32_t thr_conv, const orc_unit2 *u2, uint32_t thr_impr,
                    double *cost_seq, double *rev_seq)
{
    /* One call of simulate_epoch_of_bidding (bidding_simulation.py:44-120). */
    const double bid = (double)bid_cents / 100.0;
    const int multi_kw = kw->kind == ORC_IMPLICIT_MULTI;
    const int explicit_kw = kw->kind != ORC_IMPLICIT; /* un-rounded f64 costs: explicit and multi-bidder keywords */
    double stack_cost[256]; uint8_t stack_clicked[256]; uint32_t stack_aw[256];
    int64_t cap = n > 0 ? n : 1;
    double *slot_cost = stack_cost; uint8_t *clicked = stack_clicked; uint32_t *slot_w2 = stack_aw;
    if (cap > 256) {
        slot_cost = (double *)malloc(sizeof(double) * cap);
        clicked = (uint8_t *)malloc(cap);
        slot_w2 = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    }
    int64_t slots = 0, I = 0;
    orc_record *rec = src->rec;
    const orc_tape *tp = src->tape;
    uint32_t w[4];

    /* --- keyword.auction (classes:520-538 explicit / :623-646 + helpers:116-180) --- */
    if (!explicit_kw) {
        uint32_t R[32]; int64_t have_g = -1;
        for (int64_t a = 0; a < n; ++a) {
            int64_t j = cur->auction + a;
            int32_t c;
            int click_bit = 0, conv_bit = 0;
            if (src->mode == 0) {
                c = tp->comp_cents[tp->comp_off[k] + j];
            } else {
                /* O(clicks) tape function (section 2b): one uniform per auction decides win / click /
                 * conversion; only a clicked auction draws its price, from the competitor-bid law
                 * conditioned on losing to us, indexed by its rank among the day's clicked auctions */
                const int64_t g = j >> 5; const int p = (int)(j & 31);
                if (g != have_g) { group_uniforms(src->seed, src->env, src->step, src->agent, (uint32_t)k, (uint32_t)g, R); have_g = g; }
                const int won = (uint64_t)R[p] < u2->T1;
                click_bit = (uint64_t)R[p] < u2->T2;
                conv_bit = (uint64_t)R[p] < u2->T3;
                if (click_bit) {
                    const uint32_t r = (uint32_t)cur->n_clk++; /* rank among the day's clicked auctions */
                    draw4(src->seed, src->env, src->step, src->agent, (uint32_t)k, ST_COST, r >> 2, w);
                    c = orc_cost_cents(w[r & 3], u2);
                } else {
                    c = won ? 0 : u2->W; /* never looked at again: an unclicked win / a loss (what a recorded tape holds) */
                }
                if (rec && j < rec->cap_per_kw) { rec->comp_cents[(int64_t)k * rec->cap_per_kw + j] = c; rec->n_comp[k] = (int32_t)(j + 1); }
            }
            /* shared auction: the rivals' bids join the sampled competitor in other_bids; with
             * n=2, num_winners=1 only their maximum matters (helpers:156-177) */
            if (src->floor_cents && src->floor_cents[k] > c) c = src->floor_cents[k];
            /* nth_price_auction(n=2, num_winners=1): win iff bid > max(other bids)
             * (strict; searchsorted-left index must exceed n), cost = that maximum. */
            if ((src->mode == 1 ? u2->W : bid_cents) > c) {
                slot_cost[slots] = (double)c / 100.0;
                clicked[slots] = 0;
                slot_w2[slots] = (uint32_t)conv_bit;
                if (src->mode == 1) clicked[slots] = (uint8_t)click_bit | 0x80; /* bit7: decided */
                if (src->mode == 1 && rec) {
                    int64_t pos = (int64_t)k * rec->cap_per_kw + cur->n_click + slots;
                    if (cur->n_click + slots < rec->cap_per_kw) rec->u_click[pos] = click_bit ? 0.0 : 1.0;
                }
                ++slots; ++I;
            }
        }
    } else if (multi_kw) {
        /* default ImplicitKeyword (classes:623-688): bidders once per lane, m signed Laplace bids per
         * auction, nth_price_auction(n=2, num_winners=1) incl. zero padding for m < 3 (helpers:116-180) */
        int m = 0;
        if (src->mode == 0) {
            m = tp->impr[k * ORC_SUBSTEPS + t];
        } else {
            int mb = kw->max_bidders[k] > 0.0 ? (kw->max_bidders[k] < 62.0 ? (int)kw->max_bidders[k] : 62) : 0;
            uint32_t thr_part = orc_prob_threshold(clampd(kw->participation[k], 0.0, 1.0));
            for (int i = 0; i < mb; ++i) {
                if ((i & 3) == 0) draw4(src->seed, src->env, src->step, src->agent, (uint32_t)k, ST_BIDDERS, (uint32_t)(t * 16 + (i >> 2)), w);
                m += w[i & 3] <= thr_part;
            }
            if (rec) rec->impr[k * ORC_SUBSTEPS + t] = m;
        }
        for (int64_t a = 0; a < n; ++a) {
            int64_t j = cur->auction + a;
            double c = 0.0;
            uint32_t w_click = 0, w_conv = 0;
            if (src->mode == 0) {
                c = m < 1 ? 0.0 : tp->comp_f64[tp->comp_off[k] + j];
            } else {
                draw4(src->seed, src->env, src->step, src->agent, (uint32_t)k, ST_AUCTION, (uint32_t)(j * 16), w);
                w_click = w[0]; w_conv = w[1];
                for (int i = 0; i < m; ++i) {
                    int sl = i + 2;
                    if ((sl & 3) == 0) draw4(src->seed, src->env, src->step, src->agent, (uint32_t)k, ST_AUCTION, (uint32_t)(j * 16 + (sl >> 2)), w);
                    uint32_t wi = w[sl & 3];
                    float e = orc_neglog_u31(wi & 0x7FFFFFFFu);
                    double x = (double)fmaf((wi >> 31) ? -(float)kw->p2[k] : (float)kw->p2[k], e, (float)kw->p1[k]);
                    c = i == 0 ? x : (x > c ? x : c);
                }
                if (rec && j < rec->cap_per_kw) { rec->comp_f64[(int64_t)k * rec->cap_per_kw + j] = c; rec->n_comp[k] = (int32_t)(j + 1); }
            }
            if (m < 3 && !(c > 0.0)) c = 0.0; /* zero padding of the short auction */
            if (bid > c) {
                slot_cost[slots] = c;
                clicked[slots] = 0; slot_w2[slots] = w_conv;
                if (src->mode == 1) {
                    clicked[slots] = (uint8_t)(w_click <= thr_click) | 0x80;
                    if (rec && cur->n_click + slots < rec->cap_per_kw)
                        rec->u_click[(int64_t)k * rec->cap_per_kw + cur->n_click + slots] = (double)w_click * 2.3283064365386963e-10;
                }
                ++slots; ++I;
            }
        }
    } else {
        if (src->mode == 0) {
            I = tp->impr[k * ORC_SUBSTEPS + t];
            if (I > cap) { /* tape inconsistent with volume */
                if (cap > 256) { free(slot_cost); free(clicked); free(slot_w2); }
                return -2;
            }
            for (int64_t i = 0; i < I; ++i) {
                slot_cost[i] = tp->cost[tp->cost_off[k] + cur->n_cost + i];
                clicked[i] = 0; slot_w2[i] = 0;
            }
            slots = I;
        } else {
            for (int64_t a = 0; a < n; ++a) {
                int64_t j = cur->auction + a;
                draw4(src->seed, src->env, src->step, src->agent, (uint32_t)k, ST_AUCTION, (uint32_t)j, w);
                if (w[0] <= thr_impr) { /* Bernoulli(p): their sum is Binomial(n,p) */
                    slot_cost[slots] = orc_explicit_cost(w[3], bid);
                    clicked[slots] = (uint8_t)(w[1] <= thr_click) | 0x80;
                    slot_w2[slots] = w[2];
                    if (rec) {
                        if (cur->n_cost + slots < rec->cap_per_kw) rec->cost[(int64_t)k * rec->cap_per_kw + cur->n_cost + slots] = slot_cost[slots];
                        if (cur->n_click + slots < rec->cap_per_kw) rec->u_click[(int64_t)k * rec->cap_per_kw + cur->n_click + slots] = (double)w[1] * 2.3283064365386963e-10;
                    }
                    ++slots; ++I;
                }
            }
            if (rec) rec->impr[k * ORC_SUBSTEPS + t] = (int32_t)I;
        }
        cur->n_cost += I;
        if (I < 1) { /* phantom zero-cost slot (classes:514-515) */
            slot_cost[0] = 0.0; clicked[0] = 0; slot_w2[0] = 0; slots = 1;
            if (src->mode == 1) {
                draw4(src->seed, src->env, src->step, src->agent, (uint32_t)k, ST_PHANTOM, (uint32_t)t, w);
                clicked[0] = (uint8_t)(w[1] <= thr_click) | 0x80;
                slot_w2[0] = w[2];
                if (rec && cur->n_click < rec->cap_per_kw) rec->u_click[(int64_t)k * rec->cap_per_kw + cur->n_click] = (double)w[1] * 2.3283064365386963e-10;
            }
        }
    }
    cur->auction += n;

    /* --- sample_buyside_click: one flip per slot, all drawn (bsim:94-96, helpers:73-77) --- */
    if (src->mode == 0) {
        const double ctr = kw->ctr[k];
        for (int64_t i = 0; i < slots; ++i)
            clicked[i] = tp->u_click[tp->click_off[k] + cur->n_click + i] <= ctr;
    } else {
        for (int64_t i = 0; i < slots; ++i) clicked[i] &= 1;
        if (rec) rec->n_click[k] = (int32_t)((cur->n_click + slots) < rec->cap_per_kw ? (cur->n_click + slots) : rec->cap_per_kw);
    }
    cur->n_click += slots;

    /* --- serial budget walk (bsim:97-104) --- */
    double b = *remaining;
    double lane_cost_sum = 0.0; /* rust.sum_list(costs): sequential */
    int64_t B = 0, S = 0;
    double lane_rev[256]; double *rev = lane_rev;
    uint32_t acc_w2[256]; uint32_t *aw2 = acc_w2;
    if (slots > 256) { aw2 = (uint32_t *)malloc(sizeof(uint32_t) * slots); rev = (double *)malloc(sizeof(double) * slots); }
    for (int64_t i = 0; i < slots; ++i) {
        if (clicked[i]) {
            if (b >= slot_cost[i]) {
                aw2[B] = slot_w2[i];
                ++B;
                lane_cost_sum += slot_cost[i];
                *cost_seq += slot_cost[i]; /* whole-day sequential sum (env:235) */
                if (!explicit_kw) out->cost_cents[k] += (int64_t)llrint(slot_cost[i] * 100.0);
                b -= slot_cost[i];
            } else {
                break;
            }
        }
    }

    /* --- conversions (bsim:106-109) and revenues (bsim:111, helpers:66-70) --- */
    for (int64_t i = 0; i < B; ++i) {
        int conv;
        if (src->mode == 0) {
            conv = tp->u_conv[tp->conv_off[k] + cur->n_conv + i] <= kw->cvr[k];
        } else {
            if (explicit_kw) {
                conv = aw2[i] <= thr_conv;
                if (rec && cur->n_conv + i < rec->cap_per_kw) rec->u_conv[(int64_t)k * rec->cap_per_kw + cur->n_conv + i] = (double)aw2[i] * 2.3283064365386963e-10;
            } else {
                conv = (int)aw2[i]; /* the auction's own conversion bit (R_j < T3) */
                if (rec && cur->n_conv + i < rec->cap_per_kw) rec->u_conv[(int64_t)k * rec->cap_per_kw + cur->n_conv + i] = conv ? 0.0 : 1.0;
            }
        }
        S += conv;
    }
    cur->n_conv += B;
    if (src->mode == 1 && rec) rec->n_conv[k] = (int32_t)(cur->n_conv < rec->cap_per_kw ? cur->n_conv : rec->cap_per_kw);
    for (int64_t i = 0; i < S; ++i) {
        int32_t rc;
        int64_t r = cur->n_rev + i;
        if (src->mode == 0) {
            rc = tp->rev_cents[tp->rev_off[k] + r];
        } else {
            if ((r & 3) == 0 || i == 0)
                draw4(src->seed, src->env, src->step, src->agent, (uint32_t)k, ST_REVENUE, (uint32_t)(r >> 2), w);
            rc = orc_revenue_cents(w[r & 3], (float)kw->rev_mean[k], (float)kw->rev_std[k]);
            if (rec && r < rec->cap_per_kw) { rec->rev_cents[(int64_t)k * rec->cap_per_kw + r] = rc; rec->n_rev[k] = (int32_t)(r + 1); }
        }
        rev[i] = (double)rc / 100.0;
        *rev_seq += rev[i]; /* env:240 sequential whole-day sum */
        out->revenue_cents[k] += rc;
    }
    cur->n_rev += S;

    /* profit = rust.sum_array(revenues) - rust.sum_list(costs)  (bsim:117) */
    double lane_profit = orc_sum_array(rev, S) - lane_cost_sum;

    /* combine_outcomes addable fields (bsim:127,138-139) */
    out->impressions[k] += (int32_t)I;
    out->clicks[k] += (int32_t)B;
    out->conversions[k] += (int32_t)S;
    out->profit[k] += lane_profit;
    if (out->lane_I) {
        out->lane_I[t * kw->K + k] = (int32_t)I;
        out->lane_B[t * kw->K + k] = (int32_t)B;
        out->lane_S[t * kw->K + k] = (int32_t)S;
    }
    /* Array-valued budgets (what the action space's Box(shape=(1,)) yields) are decremented
     * IN PLACE by `budget -= cost` (bsim:102) because `budget` aliases the caller's
     * `remaining_budget` ndarray; the caller then subtracts the lane's costs again (bsim:225).
     * Scalar budgets (the notebooks' `"budget": 100000`) are immutable, so only bsim:225 acts. */
    if (alias) *remaining = b;
    *remaining -= lane_cost_sum; /* bsim:225 */

    if (cap > 256) { free(slot_cost); free(clicked); free(slot_w2); }
    if (slots > 256) { free(aw2); free(rev); }
    return 0;
}

static int step_common(const orc_keywords *kw, const int32_t *bid_cents, double budget,
                       int alias, draw_src *src, orc_result *out)
{
    const int32_t *win_cents = src->win_cents; /* free-running f32 tie rule: bid cents + bonus, or NULL */
    const int K = kw->K;
    int64_t *vol = (int64_t *)malloc(sizeof(int64_t) * K);
    kw_cursor *cur = (kw_cursor *)calloc(K, sizeof(kw_cursor));
    uint32_t *thr = (uint32_t *)malloc(sizeof(uint32_t) * 3 * K);
    orc_unit2 *u2 = (orc_unit2 *)calloc(K, sizeof(orc_unit2));
    double *cost_seq = (double *)calloc(K, sizeof(double));
    double *rev_seq = (double *)calloc(K, sizeof(double));
    uint32_t w[4];
    int rc = 0;

    /* uniform_get_auctions_per_timestep (bsim:151-167): one volume draw per keyword */
    for (int k = 0; k < K; ++k) {
        if (src->mode == 0) {
            vol[k] = src->tape->volume[k];
        } else {
            draw4(src->seed, src->env, src->step, src->agent, (uint32_t)k, ST_UNIT, 0u, w);
            vol[k] = orc_volume(w[0], kw->vol_mean[k], kw->vol_std[k]);
            if (src->rec) {
                src->rec->volume[k] = (int32_t)vol[k];
                src->rec->n_comp[k] = src->rec->n_click[k] = src->rec->n_conv[k] = 0;
                src->rec->n_rev[k] = src->rec->n_cost[k] = 0;
            }
        }
        thr[3 * k + 0] = orc_prob_threshold(kw->ctr[k]);
        thr[3 * k + 1] = orc_prob_threshold(kw->cvr[k]);
        thr[3 * k + 2] = 0;
        if (kw->kind == ORC_IMPLICIT && src->mode == 1)
            orc_unit2_make(kw->p1[k], kw->p2[k], kw->ctr[k], kw->cvr[k], win_cents ? win_cents[k] : bid_cents[k], &u2[k]);
        if (kw->kind == ORC_EXPLICIT) {
            double bid = (double)bid_cents[k] / 100.0;
            double p = orc_threshold_sigmoid(bid, kw->impression_thresh, kw->p1[k], kw->p2[k]);
            thr[3 * k + 2] = orc_prob_threshold(p);
        }
        out->impressions[k] = out->clicks[k] = out->conversions[k] = 0;
        out->cost[k] = out->revenue[k] = out->profit[k] = 0.0;
        out->cost_cents[k] = out->revenue_cents[k] = 0;
    }
    if (out->lane_I) {
        memset(out->lane_I, 0, sizeof(int32_t) * ORC_SUBSTEPS * K);
        memset(out->lane_B, 0, sizeof(int32_t) * ORC_SUBSTEPS * K);
        memset(out->lane_S, 0, sizeof(int32_t) * ORC_SUBSTEPS * K);
    }

    double remaining = budget; /* bsim:214 */
    int lanes = 0, stop = 0;
    for (int t = 0; t < ORC_SUBSTEPS && !stop; ++t) {
        for (int k = 0; k < K; ++k) {
            int64_t q = vol[k] / ORC_SUBSTEPS;
            int64_t n = (t == 0) ? vol[k] - (ORC_SUBSTEPS - 1) * q : q;
            rc = lane_run(kw, k, t, bid_cents[k], &remaining, alias, n, src, &cur[k], out,
                          thr[3 * k], thr[3 * k + 1], &u2[k], thr[3 * k + 2], &cost_seq[k], &rev_seq[k]);
            if (rc) { stop = 1; break; }
            ++lanes;
            if (remaining <= 0) { stop = 1; break; } /* bsim:230-233 */
        }
    }
    for (int k = 0; k < K; ++k) {
        out->cost[k] = cost_seq[k];
        out->revenue[k] = rev_seq[k];
        if (src->mode == 1 && src->rec) src->rec->n_cost[k] = (int32_t)cur[k].n_cost;
    }
    /* profits = rust.sum_list([kw["profit"] ...])  (env:222) */
    double reward = 0.0;
    for (int k = 0; k < K; ++k) reward += out->profit[k];
    out->reward = reward;
    out->remaining_budget = remaining;
    out->lanes_run = lanes;
    free(vol); free(cur); free(thr); free(u2); free(cost_seq); free(rev_seq);
    return rc;
}

int orc_step_replay(const orc_keywords *kw, const int32_t *bid_cents, double budget,
                    int budget_alias, const orc_tape *tape, orc_result *out)
{
    draw_src s; memset(&s, 0, sizeof s);
    s.mode = 0; s.tape = tape;
    return step_common(kw, bid_cents, budget, budget_alias, &s, out);
}

int orc_step_philox(const orc_keywords *kw, const int32_t *bid_cents, double budget,
                    int budget_alias, uint64_t seed, uint32_t env_id, uint32_t step,
                    uint32_t agent, orc_result *out, orc_record *rec)
{
    draw_src s; memset(&s, 0, sizeof s);
    s.mode = 1; s.seed = seed; s.env = env_id; s.step = step; s.agent = agent; s.rec = rec;
    return step_common(kw, bid_cents, budget, budget_alias, &s, out);
}

int orc_step_philox_shared(const orc_keywords *kw, const int32_t *bid_cents, const int32_t *floor_cents,
                           double budget, int budget_alias, uint64_t seed, uint32_t world_id, uint32_t step,
                           orc_result *out)
{   /* one bidder of a shared-auction world: every draw is keyed by the world, the clearing price
     * is max(sampled competitor, highest rival bid) */
    draw_src s; memset(&s, 0, sizeof s);
    s.mode = 1; s.seed = seed; s.env = world_id; s.step = step; s.agent = 0u; s.floor_cents = floor_cents;
    return step_common(kw, bid_cents, budget, budget_alias, &s, out);
}

/* ------------------------------------------------------------------------- */
/* 4. drift + batched driver                                                   */
/* ------------------------------------------------------------------------- */
void orc_drift_apply(int32_t K, const uint8_t *mask, int32_t num_updates, const double *coeff,
                     const double *init_std, double *vol_mean, double *ctr, double *cvr)
{   /* gymnasium_kw_env.py:132-158; zip() stops at the shortest iterable = num_updates */
    int32_t lim = num_updates < K ? num_updates : K;
    for (int32_t k = 0; k < lim; ++k) {
        if (!mask[k]) continue;
        double v = vol_mean[k] + coeff[0 * K + k] * init_std[k];
        vol_mean[k] = v > 0.0 ? v : 0.0;                 /* nonnegify */
        ctr[k] = clampd(ctr[k] * (1.0 + coeff[1 * K + k]), 0.0, 1.0); /* probify */
        cvr[k] = clampd(cvr[k] * (1.0 + coeff[2 * K + k]), 0.0, 1.0);
    }
}

void orc_drift_philox(int32_t K, uint64_t seed, uint32_t env_id, uint32_t step,
                      const double mag[3], double *coeff)
{
    uint32_t w[4];
    for (int32_t k = 0; k < K; ++k) {
        draw4(seed, env_id, step, 0u, (uint32_t)k, ST_UNIT, 0u, w);
        for (int c = 0; c < 3; ++c) {
            double u = ((double)w[1 + c] + 0.5) * 2.3283064365386963e-10;
            coeff[c * K + k] = mag[c] * (2.0 * u - 1.0);
        }
    }
}

int orc_batch_step(orc_batch *b, const double *bids, int32_t *impressions, int32_t *clicks,
                   int32_t *conversions, double *cost, double *revenue, double *reward,
                   uint8_t *terminated, uint8_t *truncated, int n_threads)
{
    const int K = b->K;
    int err = 0;
    int32_t num_updates = 0;
    if (b->drift_mask) for (int k = 0; k < K; ++k) num_updates += b->drift_mask[k] ? 1 : 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int e = 0; e < b->E; ++e) {
        int64_t po = (int64_t)e * b->param_env_stride;
        orc_keywords kw;
        kw.kind = b->kind; kw.K = K;
        kw.vol_mean = b->vol_mean + po; kw.vol_std = b->vol_std + po;
        kw.p1 = b->p1 + po; kw.p2 = b->p2 + po; kw.ctr = b->ctr + po; kw.cvr = b->cvr + po;
        kw.rev_mean = b->rev_mean + po; kw.rev_std = b->rev_std + po;
        kw.max_bidders = b->max_bidders ? b->max_bidders + po : 0;
        kw.participation = b->participation ? b->participation + po : 0;
        kw.impression_thresh = b->impression_thresh;
        int32_t *bc = (int32_t *)malloc(sizeof(int32_t) * K);
        double *profit = (double *)malloc(sizeof(double) * K);
        int64_t *cc = (int64_t *)malloc(sizeof(int64_t) * 2 * K);
        for (int k = 0; k < K; ++k) bc[k] = orc_bid_to_cents(bids[(int64_t)e * K + k]);
        orc_result r; memset(&r, 0, sizeof r);
        r.impressions = impressions + (int64_t)e * K; r.clicks = clicks + (int64_t)e * K;
        r.conversions = conversions + (int64_t)e * K;
        r.cost = cost + (int64_t)e * K; r.revenue = revenue + (int64_t)e * K;
        r.profit = profit; r.cost_cents = cc; r.revenue_cents = cc + K;
        int rc = orc_step_philox(&kw, bc, b->budget[e], b->budget_alias, b->seed, b->env_base + (uint32_t)e, b->step, 0u, &r, 0);
        if (rc) err = rc;
        reward[e] = r.reward;
        /* env tail (gymnasium_kw_env.py:222-230) */
        b->cum_profit[e] += r.reward;
        truncated[e] = b->cum_profit[e] < -b->loss_threshold;
        b->day[e] += 1;
        terminated[e] = b->day[e] >= b->max_days;
        if (b->drift_mask && b->param_env_stride) {
            double *coeff = (double *)malloc(sizeof(double) * 3 * K);
            orc_drift_philox(K, b->seed, b->env_base + (uint32_t)e, b->step, b->drift_mag, coeff);
            orc_drift_apply(K, b->drift_mask, num_updates, coeff, b->vol_std + po,
                            b->vol_mean + po, b->ctr + po, b->cvr + po);
            free(coeff);
        }
        if (terminated[e] || truncated[e]) { b->cum_profit[e] = 0.0; b->day[e] = 0; } /* auto-reset */
        free(bc); free(profit); free(cc);
    }
    b->step += 1;
    return err;
}
