// Kernels of the B200-native BiddingSimulation.step (sm_100a).
//
// Path (reference file:line):
//   BiddingSimulation.step                      adcraft/gymnasium_kw_env.py:160-269
//   simulate_epoch_of_bidding_on_campaign       adcraft/bidding_simulation.py:170-234
//   simulate_epoch_of_bidding (one "lane")      adcraft/bidding_simulation.py:44-120
//   ImplicitKeyword.auction / nth_price_auction adcraft/synthetic_kw_classes.py:623-646,
//                                               adcraft/synthetic_kw_helpers.py:116-180
//   ExplicitKeyword.auction                     adcraft/synthetic_kw_classes.py:493-538
//   update_keywords                             adcraft/gymnasium_kw_env.py:114-158
//
// Three kernels:
//   adc_lanes_philox_implicit_kernel<L>  hot kernel: L threads share one (env,keyword) unit and
//       split its day of auctions; every auction is one Philox4x32-10 call (competitor bid,
//       click and conversion words travel together), outcomes are integer cents, the unit is
//       reduced with warp shuffles and the env with L2 atomics; the thread group that
//       completes an env's last unit runs the env tail (reward, flags, auto-reset, drift).
//   adc_units_kernel<Src, kExplicit>     one thread per unit, lanes walked in order without the
//       budget (explicit keywords in free-running mode, and replay of either kind).
//   adc_serial_kernel<Src>               exact serial walk in (sub-step, keyword, click) order
//       with the shared budget, early break and the ndarray-aliasing double charge; one thread
//       per env, only for envs whose day's spend could reach the budget.
// A unit's outcome does not depend on the sub-step split unless the budget binds (implicit) --
// that is what lets the hot kernel treat the day as one flat loop ("fast path").
#include "adc_rng.cuh"
#include "adc_step.h"

#include <algorithm>
#include <cstddef>
#include <cstdio>

namespace adc {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double load_f(const void *p, int dtype, int64_t i)
{
    return dtype == ADC_F64 ? reinterpret_cast<const double *>(p)[i]
                            : (double)reinterpret_cast<const float *>(p)[i];
}

__device__ __forceinline__ void store_f(void *p, int dtype, int64_t i, double v)
{
    if (dtype == ADC_F64)
        reinterpret_cast<double *>(p)[i] = v;
    else
        reinterpret_cast<float *>(p)[i] = (float)v;
}

// Optional flat observation row of env e (adc_step_out.flat_obs): the reference's FlatArrayWrapper
// layout, keys sorted (wrappers/flat_array.py:44-87, gymnasium_kw_utils.py:383-390):
//   [buyside_clicks(K) | cost(K) | cumulative_profit | days_passed | impressions(K) | revenue(K) |
//    sellside_conversions(K)], 5 K + 2 values of float_dtype.
__device__ __forceinline__ void store_flat_unit(const adc_step_args &a, int e, int k, int I, int B, int S, double cost,
                                                double rev)
{
    if (a.out.flat_obs == nullptr) return;
    const int64_t K = a.kw.K, base = (int64_t)e * (5 * K + 2);
    store_f(a.out.flat_obs, a.out.float_dtype, base + k, (double)B);
    store_f(a.out.flat_obs, a.out.float_dtype, base + K + k, cost);
    store_f(a.out.flat_obs, a.out.float_dtype, base + 2 * K + 2 + k, (double)I);
    store_f(a.out.flat_obs, a.out.float_dtype, base + 3 * K + 2 + k, rev);
    store_f(a.out.flat_obs, a.out.float_dtype, base + 4 * K + 2 + k, (double)S);
}

// ------------------------------------------------------------------------------------------
// compact observation rows (adc_step_out.rows / adc_step_host; layout in include/adcraft_b200.h)
// ------------------------------------------------------------------------------------------
struct RowLayout {
    int64_t counts, money, tail, bytes;  // byte offsets inside a row, row size
};

__host__ __device__ inline RowLayout row_layout(int K, int float_dtype)
{
    RowLayout r;
    r.counts = 0;
    r.money = ((int64_t)6 * K + 7) & ~(int64_t)7;
    const int64_t fb = float_dtype == ADC_F64 ? 8 : 4;
    r.tail = (r.money + 2 * fb * K + 7) & ~(int64_t)7;
    r.bytes = r.tail + 24;
    return r;
}

// One warp packs env e's observation into its row: counts as uint16 (four keywords per 8-byte
// store where the row allows), money in float_dtype, the env scalars.  Reads the step's outputs
// with L2 loads: other warps wrote them before the env's completion counter let this warp in.
__device__ __forceinline__ void pack_row_warp(const adc_step_args &a, int e, unsigned char *rows, int lane)
{
    const int K = a.kw.K;
    const RowLayout L = row_layout(K, a.out.float_dtype);
    unsigned char *row = rows + (int64_t)e * L.bytes;
    const int64_t u0 = (int64_t)e * K;
    bool over = false;
    auto c16 = [&](const int32_t *p, int64_t u) {
        const int v = __ldcg(p + u);
        over = over || v > 65535;
        return (uint32_t)min(v, 65535);
    };
    const int32_t *src[3] = {a.out.impressions, a.out.clicks, a.out.conversions};
    if ((K & 3) == 0 && (L.bytes & 7) == 0) {
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            uint2 *dst = reinterpret_cast<uint2 *>(row + L.counts + (int64_t)f * 2 * K);
            for (int k4 = lane; 4 * k4 < K; k4 += 32) {
                uint2 v;
                v.x = c16(src[f], u0 + 4 * k4) | (c16(src[f], u0 + 4 * k4 + 1) << 16);
                v.y = c16(src[f], u0 + 4 * k4 + 2) | (c16(src[f], u0 + 4 * k4 + 3) << 16);
                dst[k4] = v;
            }
        }
    } else {
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            uint16_t *dst = reinterpret_cast<uint16_t *>(row + L.counts) + (int64_t)f * K;
            for (int k = lane; k < K; k += 32) dst[k] = (uint16_t)c16(src[f], u0 + k);
        }
    }
    if (a.out.float_dtype == ADC_F64) {
        double *m = reinterpret_cast<double *>(row + L.money);
        for (int k = lane; k < K; k += 32) {
            m[k] = __ldcg(reinterpret_cast<const double *>(a.out.cost) + u0 + k);
            m[K + k] = __ldcg(reinterpret_cast<const double *>(a.out.revenue) + u0 + k);
        }
    } else {
        float *m = reinterpret_cast<float *>(row + L.money);
        for (int k = lane; k < K; k += 32) {
            m[k] = __ldcg(reinterpret_cast<const float *>(a.out.cost) + u0 + k);
            m[K + k] = __ldcg(reinterpret_cast<const float *>(a.out.revenue) + u0 + k);
        }
    }
    over = __any_sync(0xFFFFFFFFu, over);
    if (lane == 0) {
        double *t = reinterpret_cast<double *>(row + L.tail);
        t[0] = __ldcg(a.out.reward + e);
        t[1] = __ldcg(a.out.obs_cum_profit + e);
        *reinterpret_cast<int32_t *>(row + L.tail + 16) = __ldcg(a.out.obs_days + e);
        const uint32_t flags = (uint32_t)__ldcg(a.out.terminated + e) | ((uint32_t)__ldcg(a.out.truncated + e) << 8) |
                               (over ? 1u << 16 : 0u);
        *reinterpret_cast<uint32_t *>(row + L.tail + 20) = flags;
    }
}

// The fast path writes a unit's part of its env's row with the unit's other outputs (spread over all
// warps, no gather) and the env scalars when the env completes; counts are below 2^16 there.
__device__ __forceinline__ void pack_row_unit(const adc_step_args &a, int e, int k, int I, int B, int S, double cost,
                                              double rev)
{
    const int K = a.kw.K;
    const RowLayout L = row_layout(K, a.out.float_dtype);
    unsigned char *row = reinterpret_cast<unsigned char *>(a.out.rows) + (int64_t)e * L.bytes;
    uint16_t *cnt = reinterpret_cast<uint16_t *>(row + L.counts);
    cnt[k] = (uint16_t)I;
    cnt[K + k] = (uint16_t)B;
    cnt[2 * K + k] = (uint16_t)S;
    if (a.out.float_dtype == ADC_F64) {
        double *m = reinterpret_cast<double *>(row + L.money);
        m[k] = cost; m[K + k] = rev;
    } else {
        float *m = reinterpret_cast<float *>(row + L.money);
        m[k] = (float)cost; m[K + k] = (float)rev;
    }
}

__device__ __forceinline__ void pack_row_tail(const adc_step_args &a, int e)
{
    const RowLayout L = row_layout(a.kw.K, a.out.float_dtype);
    unsigned char *row = reinterpret_cast<unsigned char *>(a.out.rows) + (int64_t)e * L.bytes;
    double *t = reinterpret_cast<double *>(row + L.tail);
    t[0] = a.out.reward[e];
    t[1] = a.out.obs_cum_profit[e];
    *reinterpret_cast<int32_t *>(row + L.tail + 16) = a.out.obs_days[e];
    *reinterpret_cast<uint32_t *>(row + L.tail + 20) = (uint32_t)a.out.terminated[e] | ((uint32_t)a.out.truncated[e] << 8);
}

// ------------------------------------------------------------------------------------------
// unit records (adc_step_out.unit_records): a unit's observation as one aligned 16-byte store
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_unit_record(const adc_step_args &a, int64_t u, int I, int B, int S, double cost,
                                                  double rev)
{
    const unsigned over = (I > 65535 || B > 65535 || S > 65535) ? 1u : 0u;
    uint4 r;
    r.x = (unsigned)min(I, 65535) | ((unsigned)min(B, 65535) << 16);
    r.y = (unsigned)min(S, 65535) | (over << 16);
    r.z = __float_as_uint((float)cost);
    r.w = __float_as_uint((float)rev);
    reinterpret_cast<uint4 *>(a.out.unit_records)[u] = r;
}

// env e's records from the step's finished outputs (the exact walk and the kernel families that do
// not write records themselves); `lane` of 32
__device__ __forceinline__ void pack_units_warp(const adc_step_args &a, int e, int lane)
{
    const int K = a.kw.K;
    const int64_t u0 = (int64_t)e * K;
    for (int k = lane; k < K; k += 32) {
        const int64_t u = u0 + k;
        store_unit_record(a, u, __ldcg(a.out.impressions + u), __ldcg(a.out.clicks + u), __ldcg(a.out.conversions + u),
                          (double)__ldcg(reinterpret_cast<const float *>(a.out.cost) + u),
                          (double)__ldcg(reinterpret_cast<const float *>(a.out.revenue) + u));
    }
}

// cents / 100 correctly rounded (== np.around(x, 2) of the same cents value).  The f64 division is a
// ~35-instruction routine; for |c| < 2^31 one Newton step on c * 0.01 with the exact FMA residual gives
// the identical double (checked exhaustively against c / 100.0 for every c in [0, 2^31) on the host,
// and the operations are sign-symmetric), in three FMA-pipe instructions.
__device__ __forceinline__ double cents32_to_dollars(int c)  // the same for 32-bit cents (one I2F)
{
    const double x = (double)c;
    const double q0 = __dmul_rn(x, 0.01);
    const double r = __fma_rn(-q0, 100.0, x);
    return __fma_rn(r, 0.01, q0);
}

__device__ __forceinline__ double cents_to_dollars(long long c)
{
    const double x = (double)c;
    if (c >= (1LL << 31) || c <= -(1LL << 31)) return __ddiv_rn(x, 100.0);
    const double q0 = __dmul_rn(x, 0.01);
    const double r = __fma_rn(-q0, 100.0, x);
    return __fma_rn(r, 0.01, q0);
}

// budget of env e for this step: action budget rounded to cents (env:199) or the persisted one
__device__ __forceinline__ double step_budget(const adc_step_args &a, int e)
{
    if (a.budget_in != nullptr) {
        const double b = load_f(a.budget_in, a.bids_dtype, e);
        return __ddiv_rn(rint(__dmul_rn(b, 100.0)), 100.0);
    }
    return a.env.budget[e];
}

// Can the day's total spend possibly make a `budget >= cost` check fail or drive the remaining
// budget to <= 0?  `spend` is the exact (cents) or f64 total; the margin covers the rounding of
// the reference's sequential float subtractions (<= n_clicks * ulp(budget)).
__device__ __forceinline__ bool budget_is_safe(double budget, double spend, int alias)
{
    const double s = alias ? 2.0 * spend : spend;
    return s + 0.005 + 1e-7 * fabs(budget) + 1e-9 * s < budget;
}

// The serial queue length and the hot kernel's work counter are double-buffered on the step
// parity: every first-stage kernel of a step zeroes the copies the NEXT step will use, so no
// reset launch is needed (the previous step, the last user of those copies, has completed).
__device__ __forceinline__ void reset_next_counters(const adc_step_args &a)
{
    a.scratch.serial_count[(a.parity & 1u) ^ 1u] = 0;
    if (a.scratch.work_counter != nullptr) a.scratch.work_counter[(a.parity & 1u) ^ 1u] = 0u;
}

struct Drift3 {
    double c[3];
};

__device__ __forceinline__ Drift3 drift_from_words(const adc_step_args &a, uint4 w)
{
    Drift3 d;
    const uint32_t ws[3] = {w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double u = __dmul_rn(__dadd_rn((double)ws[i], 0.5), 2.3283064365386963e-10);
        d.c[i] = __dmul_rn(a.drift.mag[i], __dsub_rn(__dmul_rn(2.0, u), 1.0));
    }
    return d;
}

// update_keywords for one keyword (env:145-158): nonnegify / probify.  The four loads are issued
// before the first store (one DRAM round trip per keyword instead of three dependent ones).
struct DriftState {
    double vol_mean, vol_std, ctr, cvr;
};

__device__ __forceinline__ DriftState drift_load(const adc_step_args &a, int e, int k)
{
    const int64_t i = (int64_t)e * a.kw.env_stride + k;
    DriftState s;
    s.vol_mean = a.kw.vol_mean[i];
    s.vol_std = a.kw.vol_std[i];
    s.ctr = a.kw.ctr[i];
    s.cvr = a.kw.cvr[i];
    return s;
}

__device__ __forceinline__ void drift_store(const adc_step_args &a, int e, int k, const DriftState &s, const Drift3 &d)
{
    const int64_t i = (int64_t)e * a.kw.env_stride + k;
    const double vm = __dadd_rn(s.vol_mean, __dmul_rn(d.c[0], s.vol_std));
    a.kw.vol_mean[i] = vm > 0.0 ? vm : 0.0;
    a.kw.ctr[i] = clampd(__dmul_rn(s.ctr, __dadd_rn(1.0, d.c[1])), 0.0, 1.0);
    a.kw.cvr[i] = clampd(__dmul_rn(s.cvr, __dadd_rn(1.0, d.c[2])), 0.0, 1.0);
}

__device__ __forceinline__ void drift_apply(const adc_step_args &a, int e, int k, const Drift3 &d)
{
    drift_store(a, e, k, drift_load(a, e, k), d);
}

__device__ __forceinline__ bool drift_wanted(const adc_step_args &a, int k)
{
    return a.drift.mask != nullptr && k < a.drift.num_updates && a.drift.mask[k] != 0;
}

// env tail (env:222-230) + auto-reset.  reward in dollars.
__device__ __forceinline__ void env_tail(const adc_step_args &a, int e, double reward,
                                         double budget, double remaining)
{
    const double cum = __dadd_rn(a.env.cum_profit[e], reward);
    const bool trunc = cum < -a.env.loss_threshold;
    const int day = a.env.day[e] + 1;
    const bool term = day >= a.env.max_days;
    a.out.reward[e] = reward;
    a.out.obs_cum_profit[e] = cum;
    a.out.obs_days[e] = day;
    if (a.out.flat_obs != nullptr) {
        const int64_t K = a.kw.K, base = (int64_t)e * (5 * K + 2);
        store_f(a.out.flat_obs, a.out.float_dtype, base + 2 * K, cum);
        store_f(a.out.flat_obs, a.out.float_dtype, base + 2 * K + 1, (double)day);
    }
    a.out.terminated[e] = term ? 1 : 0;
    a.out.truncated[e] = trunc ? 1 : 0;
    if (a.out.remaining_budget) a.out.remaining_budget[e] = remaining;
    if (a.out.episode_reward != nullptr) a.out.episode_reward[e] = __dadd_rn(a.out.episode_reward[e], reward);
    if (a.out.episode_count != nullptr && (term || trunc)) a.out.episode_count[e] += 1;
    const bool done = a.autoreset && (term || trunc);
    a.env.cum_profit[e] = done ? 0.0 : cum;
    a.env.day[e] = done ? 0 : day;
    // an ndarray budget is mutated in place by the lanes, so the leftover persists (bsim:102)
    a.env.budget[e] = a.budget_alias ? remaining : budget;
}

// ------------------------------------------------------------------------------------------
// draw sources
// ------------------------------------------------------------------------------------------
struct PhiloxSrc {
    static constexpr bool kTape = false;
    uint32_t k0, k1, step, env;
    __device__ __forceinline__ uint4 draw(uint32_t stream, uint32_t kw, uint32_t idx) const
    {
        return philox4x32_10(idx, step, stream_word(stream, 0u, kw), env, k0, k1);
    }
};

struct TapeSrc {
    static constexpr bool kTape = true;
    const adc_tape *t;
};

// running cursors of one unit inside one env step
struct UnitCur {
    long long auction;  // auctions evaluated so far (day-level ordinal)
    int n_click;        // click slots drawn so far
    int n_conv;         // accepted clicks so far
    int n_rev;          // conversions so far
    int n_cost;         // explicit: impressions so far
    int n_clk;          // free-running implicit: clicked auctions so far, accepted or not (price draw rank)
};

struct UnitPar {
    int bid_cents;
    int win_cents;  // what the competitor bid is compared with: bid_cents (+1 under the f32 tie rule)
    double bid;  // dollars (explicit)
    float loc, scale, rev_mean, rev_sd;
    double ctr, cvr;
    uint32_t thr_click, thr_conv, thr_impr;
    int floor_cents;  // shared auctions: highest rival bid (INT_MIN when there are no rivals)
    int multi;        // ADC_IMPLICIT_MULTI: m bidders per lane, signed un-rounded Laplace bids (classes:649-688)
    int max_bidders;
    uint32_t thr_part;
    Unit2 u2;         // free-running implicit keywords: thresholds + price sampler (adc_rng.cuh)
    long long volume; // free-running implicit keywords: the day's volume (group masks need it)
};

// cents the win test uses (adc_step_args.f32_ties)
__device__ __forceinline__ int win_cents_of(const adc_step_args &a, int bid_cents)
{
    return bid_cents + ((a.f32_ties && a.bids_dtype == ADC_F32) ? bid_tie_bonus(bid_cents) : 0);
}

// Shared auctions: the highest RIVAL bid of unit (e, k) in cents (INT_MIN without rivals).  Either the
// caller's floor_cents table or, when env_group = A > 1 comes without one, computed here from the A
// bid rows of the world: the unique top bidder faces the runner-up, everybody else (tied leaders
// too) faces the top bid -- nth_price_auction(n=2, num_winners=1) on the rivals (helpers:116-180).
__device__ __forceinline__ int unit_floor(const adc_step_args &a, int e, int k, int own_cents)
{
    if (a.floor_cents != nullptr) return a.floor_cents[(int64_t)e * a.kw.K + k];
    if (a.env_group <= 1) return (int)0x80000000;
    const int A = a.env_group, w0 = (e / A) * A;
    int top = (int)0x80000000, second = (int)0x80000000, n_top = 0;
    for (int r = 0; r < A; ++r) {
        const int c = bid_to_cents(load_f(a.bids, a.bids_dtype, (int64_t)(w0 + r) * a.kw.K + k));
        if (c > top) { second = top; top = c; n_top = 1; }
        else if (c == top) ++n_top;
        else if (c > second) second = c;
    }
    return (own_cents == top && n_top == 1) ? second : top;
}

// Philox env id: the A bidders of a shared-auction world draw from the same counters.
__device__ __forceinline__ uint32_t philox_env(const adc_step_args &a, int e)
{
    return a.env_base + (uint32_t)(a.env_group > 1 ? e / a.env_group : e);
}

struct LaneOut {
    int I, B, S;
    long long cost_cents, rev_cents;
    double lane_cost_sum;  // dollars, sequential rust.sum_list of this lane's costs (bsim:225)
    bool overrun;          // tape mode: a stream ended before the walk did (see tape_at)
};

// Bounds-checked tape read.  A tape recorded from a run whose budget bound holds only what that
// run consumed (lanes after the early break drew nothing), so the budget-free walk can run off
// the end of a unit's stream: it then reads a harmless default, reports `overrun`, and the env
// is sent to the exact serial kernel, which consumes precisely what the reference consumed.
template <typename T>
__device__ __forceinline__ T tape_at(const T *vals, const int64_t *off, int64_t u, int64_t i, T dflt,
                                     bool &overrun)
{
    const int64_t idx = off[u] + i;
    if (idx >= off[u + 1]) {
        overrun = true;
        return dflt;
    }
    return vals[idx];
}

// Free-running implicit keywords: clicked auctions among the day's first j0 (the price-draw rank of
// the next click).  The generic thread-serial walk recounts it per lane; the warp kernel keeps it in
// its slab.
__device__ __noinline__ int clicks_before(const PhiloxSrc &src, int kw, const Unit2 &u2, long long volume, long long j0)
{
    const PhiloxPre pa = philox_pre(src.step, stream_word(ST_AUCTION, 0u, (uint32_t)kw), src.env, src.k0, src.k1);
    int n = 0;
    for (long long gbase = 0; gbase < j0; gbase += 32) {
        const long long vrem = volume - gbase;
        const uint32_t active = vrem >= 32 ? 0xFFFFFFFFu : ((1u << (int)vrem) - 1u);
        const Masks3 m = group_masks(active, (uint32_t)(gbase >> 5), u2.t1, u2.t2, u2.t3, u2.full, pa.n0, pa.n1, pa.x3,
                                     src.k0, src.k1);
        const long long lim = j0 - gbase;
        n += __popc(m.click & (lim >= 32 ? 0xFFFFFFFFu : ((1u << (int)lim) - 1u)));
    }
    return n;
}

// One call of simulate_epoch_of_bidding (bsim:44-120) for unit u, sub-step t, n auctions.
// kBudget=false: budget is +inf (no check can fail).  `b` is the lane-local budget copy;
// `day_cost` is the keyword's running sequential f64 cost sum over the day's concatenated
// clicks (env:235) -- only the explicit keywords' un-rounded costs need it.
template <typename Src, bool kExplicit, bool kBudget>
__device__ __forceinline__ LaneOut lane_walk(const Src &src, const adc_tape *tp, int64_t u, int kw,
                                             int t, long long n, const UnitPar &p, UnitCur &cur,
                                             double &b, double &day_cost,
                                             const adc_detail *det = nullptr)
{
    LaneOut o;
    o.I = o.B = o.S = 0;
    o.cost_cents = o.rev_cents = 0;
    o.lane_cost_sum = 0.0;
    o.overrun = false;
    bool stopped = false;  // the `break` of bsim:103-104: later slots are drawn but not scanned
    int slots = 0;
    uint4 rw = make_uint4(0, 0, 0, 0);
    bool rw_valid = false;

    auto on_slot = [&](double cost, long long cost_c, bool clicked, uint32_t w2) {
        // clicked slot: budget walk (bsim:97-104), conversion flip (bsim:106-109), revenue (bsim:111)
        if (!clicked || stopped) return;
        if (kBudget && !(b >= cost)) {
            stopped = true;
            return;
        }
        bool conv;
        if constexpr (Src::kTape)
            conv = tape_at(tp->u_conv, tp->conv_off, u, cur.n_conv + o.B, 2.0, o.overrun) <= p.cvr;
        else if constexpr (kExplicit)
            conv = w2 <= p.thr_conv;
        else
            conv = w2 != 0u;  // the auction's own conversion bit (R_j < T3)
        o.B += 1;
        o.cost_cents += cost_c;
        if constexpr (kExplicit) day_cost = __dadd_rn(day_cost, cost);
        o.lane_cost_sum = __dadd_rn(o.lane_cost_sum, cost);
        if (kBudget) b = __dsub_rn(b, cost);
        const int click_ord = cur.n_conv + o.B - 1;  // ordinal of this accepted click in the day
        const bool rec = det != nullptr && det->costs != nullptr && click_ord < det->cap;
        if (rec) {
            det->costs[u * det->cap + click_ord] = cost;
            det->rev_per_cost[u * det->cap + click_ord] = 0.0;
        }
        if (conv) {
            const int r = cur.n_rev + o.S;
            int rc;
            if constexpr (Src::kTape) {
                rc = tape_at(tp->rev_cents, tp->rev_off, u, r, 0, o.overrun);
            } else {
                if (!rw_valid || (r & 3) == 0) {
                    rw = src.draw(ST_REVENUE, (uint32_t)kw, (uint32_t)(r >> 2));
                    rw_valid = true;
                }
                const uint32_t w = (r & 3) == 0 ? rw.x : (r & 3) == 1 ? rw.y : (r & 3) == 2 ? rw.z : rw.w;
                rc = revenue_cents(w, p.rev_mean, p.rev_sd);
            }
            o.S += 1;
            o.rev_cents += rc;
            if (rec) det->rev_per_cost[u * det->cap + click_ord] = cents_to_dollars(rc);
        }
    };

    if constexpr (!kExplicit) {
        // second-price auction against one competitor (classes:644-646, helpers:116-180):
        // win iff bid > competitor (strict), cost = competitor's bid.
        if constexpr (Src::kTape) {
            for (long long a = 0; a < n; ++a) {
                const long long j = cur.auction + a;
                const int c = tape_at(tp->comp_cents, tp->comp_off, u, j, 0x7FFFFFFF, o.overrun);
                if (p.win_cents > c) {
                    const bool clicked =
                        tape_at(tp->u_click, tp->click_off, u, cur.n_click + slots, 2.0, o.overrun) <= p.ctr;
                    on_slot(cents_to_dollars(c), c, clicked, 0u);
                    ++slots;
                    ++o.I;
                }
            }
        } else {
            // O(clicks) tape function (adc_rng.cuh): one bit-sliced uniform per auction decides win /
            // click / conversion; a clicked auction draws its price by its click rank in the day
            const PhiloxPre pa = philox_pre(src.step, stream_word(ST_AUCTION, 0u, (uint32_t)kw), src.env, src.k0, src.k1);
            const bool beats_rivals = p.u2.W > p.floor_cents;
            const long long j_end = cur.auction + n;
            long long j = cur.auction;
            int nclk = 0;
            uint4 cw = make_uint4(0, 0, 0, 0);
            int cw_idx = -1;
            while (j < j_end) {
                const uint32_t g = (uint32_t)(j >> 5);
                const long long gbase = (long long)g << 5;
                const long long vrem = p.volume - gbase;
                const uint32_t active = vrem >= 32 ? 0xFFFFFFFFu : ((1u << (int)vrem) - 1u);
                const Masks3 m = group_masks(active, g, p.u2.t1, p.u2.t2, p.u2.t3, p.u2.full, pa.n0, pa.n1, pa.x3,
                                             src.k0, src.k1);
                const int p0 = (int)(j - gbase);
                const int p1 = (int)((j_end - gbase) < 32 ? (j_end - gbase) : 32);
                const uint32_t range = (p1 >= 32 ? 0xFFFFFFFFu : ((1u << p1) - 1u)) & ~((1u << p0) - 1u);
                uint32_t wins = m.win & range;
                while (wins) {
                    const int b = __ffs(wins) - 1;
                    wins &= wins - 1;
                    const bool clk = (m.click >> b) & 1u;
                    int c = 0;
                    if (clk) {
                        const int r = cur.n_clk + nclk;
                        ++nclk;
                        if ((r >> 2) != cw_idx) {
                            cw_idx = r >> 2;
                            cw = src.draw(ST_COST, (uint32_t)kw, (uint32_t)cw_idx);
                        }
                        const uint32_t w = (r & 3) == 0 ? cw.x : (r & 3) == 1 ? cw.y : (r & 3) == 2 ? cw.z : cw.w;
                        c = cost_cents2(w, p.u2.t1, (p.u2.full & 1u) != 0u, p.u2.h1, p.u2.a1, p.u2.a2, p.u2.L,
                                        p.u2.b, p.u2.W, kNeglogTab);
                    }
                    if (!beats_rivals) continue;  // shared auction: a rival bids at least as much
                    c = max(c, p.floor_cents);
                    on_slot(cents_to_dollars(c), c, clk, (m.conv >> b) & 1u);
                    ++slots;
                    ++o.I;
                }
                j = gbase + p1;
            }
            cur.n_clk += nclk;
        }
    } else if (p.multi) {
        // Default ImplicitKeyword (classes:623-688): m ~ Binomial(max_bidders, participation) bidders drawn
        // ONCE for the lane, every auction's other bids are m signed Laplace draws, cleared by
        // nth_price_auction(n=2, num_winners=1) (helpers:116-180): with fewer than 3 bidders zeros are
        // appended, so the clearing price is max(bids..., and 0 when m < 3); win iff bid > that (strict),
        // cost = that price, un-rounded and possibly negative.
        int m = 0;
        if constexpr (Src::kTape) {
            m = tp->impr[u * ADC_SUBSTEPS + t];  // the lane's bidder count rides in the impr array
        } else {
            uint4 w = make_uint4(0, 0, 0, 0);
            for (int i = 0; i < p.max_bidders; ++i) {
                if ((i & 3) == 0) w = src.draw(ST_BIDDERS, (uint32_t)kw, (uint32_t)(t * 16 + (i >> 2)));
                const uint32_t wi = (i & 3) == 0 ? w.x : (i & 3) == 1 ? w.y : (i & 3) == 2 ? w.z : w.w;
                m += wi <= p.thr_part;
            }
        }
        for (long long a = 0; a < n; ++a) {
            const long long j = cur.auction + a;
            double c = 0.0;
            uint32_t w_click = 0u, w_conv = 0u;
            if constexpr (Src::kTape) {
                c = tape_at(tp->comp_f64, tp->comp_off, u, j, __longlong_as_double(0x7FF0000000000000LL), o.overrun);
                if (m < 1) c = 0.0;
            } else {
                uint4 w = src.draw(ST_AUCTION, (uint32_t)kw, (uint32_t)(j * 16));
                w_click = w.x; w_conv = w.y;
                for (int i = 0; i < m; ++i) {  // bidder i's word: slot i + 2 of the auction's 64 words
                    const int sl = i + 2;
                    if ((sl & 3) == 0) w = src.draw(ST_AUCTION, (uint32_t)kw, (uint32_t)(j * 16 + (sl >> 2)));
                    const uint32_t wi = (sl & 3) == 0 ? w.x : (sl & 3) == 1 ? w.y : (sl & 3) == 2 ? w.z : w.w;
                    const double x = laplace_signed(wi, p.loc, p.scale);
                    c = i == 0 ? x : (x > c ? x : c);
                }
            }
            if (m < 3 && !(c > 0.0)) c = 0.0;  // zero padding of the short auction (helpers:156-161)
            if (p.bid > c) {
                bool clicked;
                if constexpr (Src::kTape)
                    clicked = tape_at(tp->u_click, tp->click_off, u, cur.n_click + slots, 2.0, o.overrun) <= p.ctr;
                else
                    clicked = w_click <= p.thr_click;
                on_slot(c, 0, clicked, w_conv);
                ++slots;
                ++o.I;
            }
        }
    } else {
        if constexpr (Src::kTape) {
            const int I = tp->impr[u * ADC_SUBSTEPS + t];
            for (int i = 0; i < I; ++i) {
                const double cost = tape_at(tp->cost, tp->cost_off, u, cur.n_cost + i, 0.0, o.overrun);
                const bool clicked = tape_at(tp->u_click, tp->click_off, u, cur.n_click + i, 2.0, o.overrun) <= p.ctr;
                on_slot(cost, 0, clicked, 0u);
            }
            o.I = I;
            slots = I;
            if (I < 1) {  // phantom zero-cost slot (classes:514-515)
                const bool clicked = tape_at(tp->u_click, tp->click_off, u, cur.n_click, 2.0, o.overrun) <= p.ctr;
                on_slot(0.0, 0, clicked, 0u);
                slots = 1;
            }
        } else {
            for (long long a = 0; a < n; ++a) {
                const long long j = cur.auction + a;
                const uint4 w = src.draw(ST_AUCTION, (uint32_t)kw, (uint32_t)j);
                if (w.x <= p.thr_impr) {  // Bernoulli(p); the lane's sum is Binomial(n,p)
                    on_slot(explicit_cost(w.w, p.bid), 0, w.y <= p.thr_click, w.z);
                    ++slots;
                    ++o.I;
                }
            }
            if (o.I < 1) {
                const uint4 w = src.draw(ST_PHANTOM, (uint32_t)kw, (uint32_t)t);
                on_slot(0.0, 0, w.y <= p.thr_click, w.z);
                slots = 1;
            }
        }
        cur.n_cost += o.I;
    }
    cur.auction += n;
    cur.n_click += slots;
    cur.n_conv += o.B;
    cur.n_rev += o.S;
    return o;
}

// free_running: also build the O(clicks) thresholds / price sampler of an implicit keyword
__device__ __forceinline__ UnitPar load_unit_par(const adc_step_args &a, int e, int k, bool free_running = false)
{
    UnitPar p;
    const int64_t pi = (int64_t)e * a.kw.env_stride + k;
    const int64_t u = (int64_t)e * a.kw.K + k;
    p.bid_cents = bid_to_cents(load_f(a.bids, a.bids_dtype, u));
    p.win_cents = win_cents_of(a, p.bid_cents);
    p.bid = cents_to_dollars(p.bid_cents);
    p.ctr = a.kw.ctr[pi];
    p.cvr = a.kw.cvr[pi];
    p.thr_click = prob_threshold(p.ctr);
    p.thr_conv = prob_threshold(p.cvr);
    p.rev_mean = (float)a.kw.rev_mean[pi];
    p.rev_sd = (float)a.kw.rev_std[pi];
    p.loc = (float)a.kw.p1[pi];
    p.scale = (float)a.kw.p2[pi];
    p.thr_impr = 0u;
    p.floor_cents = unit_floor(a, e, k, p.bid_cents);
    p.volume = 0;
    p.multi = 0; p.max_bidders = 0; p.thr_part = 0u;
    if (a.kw.kind == ADC_EXPLICIT) {
        p.thr_impr = prob_threshold(threshold_sigmoid(p.bid, a.kw.impression_thresh, a.kw.p1[pi], a.kw.p2[pi]));
    } else if (a.kw.kind == ADC_IMPLICIT_MULTI) {
        p.multi = 1;
        const double mb = a.kw.max_bidders[pi];
        p.max_bidders = mb > 0.0 ? (mb < 62.0 ? (int)mb : 62) : 0;  // 2 + 15 x 4 bid words per auction
        p.thr_part = prob_threshold(clampd(a.kw.participation[pi], 0.0, 1.0));  // probify (classes:663)
    } else if (free_running)
        p.u2 = unit2_make(a.kw.p1[pi], a.kw.p2[pi], p.ctr, p.cvr, p.win_cents);
    return p;
}

template <typename Src>
__device__ __forceinline__ long long unit_volume(const adc_step_args &a, const Src &src,
                                                 const adc_tape *tp, int e, int k, uint4 *unit_words)
{
    if constexpr (Src::kTape) {
        *unit_words = make_uint4(0, 0, 0, 0);
        return tp->volume[(int64_t)e * a.kw.K + k];
    } else {
        const int64_t pi = (int64_t)e * a.kw.env_stride + k;
        const uint4 w = src.draw(ST_UNIT, (uint32_t)k, 0u);
        *unit_words = w;
        long long v = volume_draw(w.x, a.kw.vol_mean[pi], a.kw.vol_std[pi]);
        return v > 2147483647LL ? 2147483647LL : v;
    }
}

template <typename Src>
__device__ __forceinline__ Drift3 unit_drift(const adc_step_args &a, const adc_tape *tp, int e, int k,
                                             uint4 unit_words)
{
    if constexpr (Src::kTape) {
        Drift3 d;
        const int K = a.kw.K;
#pragma unroll
        for (int i = 0; i < 3; ++i) d.c[i] = tp->drift ? tp->drift[((int64_t)e * 3 + i) * K + k] : 0.0;
        return d;
    } else {
        return drift_from_words(a, unit_words);
    }
}

// Unit finished: publish it, and if it was the env's last unit, run the env tail or queue the
// env for the exact serial walk.  Returns (to the calling thread only) 1 if this env is safe and
// the caller should apply the drift, 0 otherwise.  Called by ONE thread per unit.
constexpr int kEnvForceBit = 1 << 30;  // adc_scratch.env_done: some unit of the env asked for the exact walk

__device__ __forceinline__ int unit_done(const adc_step_args &a, int e, long long profit_cents,
                                         long long cost_cents, bool force = false, bool fenced = false, int n_units = 1)
{
    if (profit_cents != 0) atomicAdd(reinterpret_cast<unsigned long long *>(a.scratch.env_profit + e),
                                     (unsigned long long)profit_cents);
    if (cost_cents != 0) atomicAdd(reinterpret_cast<unsigned long long *>(a.scratch.env_cost + e),
                                   (unsigned long long)cost_cents);
    // `force`: the unit could not be evaluated here (beyond the kernel's caps, a tape that ran out, an
    // env routed by adc_scratch.serial_hint).  The flag rides in the counter word: this thread's OR
    // precedes its own increment, so the last finisher's increment sees every unit's flag.
    if (force) atomicOr(a.scratch.env_done + e, kEnvForceBit);
    if (!fenced) __threadfence();  // (`fenced`: the caller's stores are already ordered before this call)
    const int old = atomicAdd(a.scratch.env_done + e, n_units);  // (n_units > 1: a warp's units of one env, published by one lane)
    if ((old & (kEnvForceBit - 1)) != a.kw.K - n_units) return 0;
    __threadfence();
    a.scratch.env_done[e] = 0;
    const long long profit =
        (long long)atomicExch(reinterpret_cast<unsigned long long *>(a.scratch.env_profit + e), 0ull);
    const long long cost =
        (long long)atomicExch(reinterpret_cast<unsigned long long *>(a.scratch.env_cost + e), 0ull);
    const double budget = step_budget(a, e);
    double reward, spend;
    if (a.kw.kind != ADC_IMPLICIT) {
        // un-rounded costs (explicit / multi-bidder keywords): sum the per-unit f64 results in keyword order (env:222)
        reward = 0.0;
        spend = 0.0;
        const int K = a.kw.K;
        for (int k = 0; k < K; ++k) {
            const int64_t u = (int64_t)e * K + k;
            const double c = __ldcg(a.scratch.unit_cost_f64 + u);
            const double r = cents_to_dollars(__ldcg(a.out.revenue_cents + u));
            reward = __dadd_rn(reward, __dsub_rn(r, c));
            spend = __dadd_rn(spend, c);
        }
    } else {
        reward = cents_to_dollars(profit);
        spend = cents_to_dollars(cost);
    }
    if (a.force_serial || (old & kEnvForceBit) || !budget_is_safe(budget, spend, a.budget_alias)) {
        const int slot = atomicAdd(a.scratch.serial_count + (a.parity & 1u), 1);
        a.scratch.serial_list[slot] = e;
        return 0;
    }
    const double used = a.budget_alias ? __dmul_rn(2.0, spend) : spend;
    env_tail(a, e, reward, budget, __dsub_rn(budget, used));
    return 1;
}

// ------------------------------------------------------------------------------------------
// hot kernel (free-running implicit keywords, budget cannot bind): the O(clicks) step.
// A warp takes a batch of 32 units.
//   setup     lane <-> unit: parameters, volume draw, the unit's thresholds T1 >= T2 >= T3 and its
//             price sampler (three deterministic exps in f64, adc_rng.cuh unit2_make);
//   outcomes  lane <-> unit: the day's auctions in groups of 32, one bit-sliced uniform each
//             (group_masks): impressions / clicks / conversions are three popcounts per group;
//             a group costs ~4 Philox calls whatever happens inside it;
//   prices    one price per CLICK, 4 per Philox call, flattened over the batch (call i belongs to
//             the unit whose prefix range holds i) so every trip is full;
//   revenues  one per conversion, flattened the same way;
//   outputs   lane <-> unit: coalesced stores, env completion by L2 atomics, the last finisher of
//             an env runs its tail, its episode accumulation and its drift.
// Batches come from an atomic work counter when the caller provides one.
// ------------------------------------------------------------------------------------------
struct __align__(16) FlatCost {  // 48 B per unit in shared memory: the price sampler + its Philox pre-round
    uint32_t t1, h1, a1, a2;
    float L, b;
    int W;          // bit 31: T1 is 2^32
    int floor_c;
    uint32_t n0, n1, x3;
    int B;
};

struct __align__(16) FlatGroups {  // 32 B per unit in shared memory: what a (unit, group) item needs of its unit
    uint32_t t1, t2, t3, full;
    uint32_t n0, n1, x3;
    int V;
};

struct __align__(16) FlatRev {
    float mean, sd;
    int S;
    uint32_t n0;
    uint32_t n1, x3, pad0, pad1;
};

// The hot kernel's work list: pull index -> a chunk of consecutive batches.
// Big batches of 32 units first: as many as there are when the warps pull dynamically, whole
// static rounds otherwise; then the rest in small batches of kFlatTail units.
// A dynamic pull hands out up to 16 big batches (large steps: one pull per batch would be several
// 10^5 atomics on one address per millisecond) while every warp still makes >= 8 pulls.
// Kept out of line: its loop-invariant terms would otherwise occupy registers of the hot loops.
constexpr int kFlatTail = 8;

struct FlatChunk {
    int64_t u0;  // first unit of the next batch
    int cnt;     // units per batch (32 or kFlatTail)
    int left;    // batches left in the chunk (0: no more work)
};

__device__ __noinline__ FlatChunk flat_chunk(int64_t total, int64_t n_warps, bool dynamic, int64_t pi)
{
    const int64_t n_big = dynamic ? total / 32
                                  : total / (32 * n_warps) * n_warps;
    const int64_t n_small = (total - n_big * 32 + kFlatTail - 1) / kFlatTail;
    const int64_t chunk = dynamic ? min((int64_t)16, max((int64_t)1, n_big / (8 * n_warps))) : 1;
    const int64_t n_chunks = (n_big + chunk - 1) / chunk;
    FlatChunk c;
    c.u0 = 0; c.cnt = 0; c.left = 0;
    if (pi < n_chunks) {
        const int64_t b = pi * chunk;
        c.u0 = b * 32; c.cnt = 32; c.left = (int)min(chunk, n_big - b);
    } else if (pi - n_chunks < n_small) {
        c.u0 = n_big * 32 + (pi - n_chunks) * kFlatTail; c.cnt = kFlatTail; c.left = 1;
    }
    return c;
}

// 7 warps x 4 CTAs = 28 warps per SM at 72 registers (8 x 3 = 24: C2 0.188 -> 0.179 ms, C3 very sparse 1.61 -> 1.54 ms)
constexpr int kFlatWarps = 7;
// Caps of the fast kernel's 32-bit sums; a unit beyond them sends its env to the exact serial kernel
// instead (volumes and bids this large do not occur in the reference's configs).
constexpr int kMaxFlatVolume = 65535;
constexpr int kMaxFlatBidCents = 65535;
constexpr long long kMaxFlatSpendCents = 0xFFFFFFFFLL;  // volume x bid: the unit's cost sum fits two 24-bit halves

__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, v, off);
        if (lane >= off) v += t;
    }
    return v;
}

// Episode accumulation of a finished env (optional adc_step_out.episode_profit_cents): the step's
// exact profit per keyword, added once the env's step is final.  `lane` of `n_lanes` cooperating lanes.
__device__ __forceinline__ void episode_accumulate(const adc_step_args &a, int e, int lane, int n_lanes)
{
    if (a.out.episode_profit_cents == nullptr) return;
    const int K = a.kw.K;
    for (int k = lane; k < K; k += n_lanes) {
        const int64_t u = (int64_t)e * K + k;
        const long long r = __ldcg(a.out.revenue_cents + u), c = __ldcg(a.out.cost_cents + u);
        a.out.episode_profit_cents[u] += r - c;
    }
}

// Flattened work of a batch: unit `lane` has `n` items (Philox calls); item i of the batch belongs to
// the unit whose prefix range holds i.  flat_map_begin publishes, per warp, the exclusive prefix
// (`start`, shared [32]) and the list of units that have items (`nzl`, shared [32]); flat_map_unit
// then finds the unit of item base + lane without a search: every owner lane marks where its unit
// ends inside the trip, one OR-REDUX + popc ranks the lane's unit among those with items.
struct FlatMap {
    int incl;   // inclusive prefix of this lane's unit
    bool nz;    // this lane's unit has items
    int total;  // items of the whole batch
};

__device__ __forceinline__ FlatMap flat_map_begin(int n, int lane, int *start, unsigned char *nzl)
{
    FlatMap m;
    m.incl = warp_incl_scan(n, lane);
    m.nz = n > 0;
    const unsigned nzmask = __ballot_sync(0xFFFFFFFFu, m.nz);
    __syncwarp();  // earlier readers of the two tables are done
    if (m.nz) nzl[__popc(nzmask & ((1u << lane) - 1u))] = (unsigned char)lane;
    start[lane] = m.incl - n;
    __syncwarp();
    m.total = __shfl_sync(0xFFFFFFFFu, m.incl, 31);
    return m;
}

// unit of item base + lane (valid when base + lane < total); all lanes must call
__device__ __forceinline__ int flat_map_unit(const FlatMap &m, int base, int lane, const unsigned char *nzl)
{
    const int rel = m.incl - base;
    const unsigned marks = __reduce_or_sync(0xFFFFFFFFu, (m.nz && rel > 0 && rel < 32) ? (1u << rel) : 0u);
    const int before = __popc(__ballot_sync(0xFFFFFFFFu, m.nz && rel <= 0));
    return nzl[min(before + __popc(marks & (0xFFFFFFFFu >> (31 - lane))), 31)];
}

template <bool kFloor, bool kSpread>
__global__ void __launch_bounds__(kFlatWarps * 32)
adc_flat2_implicit_kernel(const __grid_constant__ adc_step_args a)
{
    __shared__ FlatCost s_cost[kFlatWarps][32];
    __shared__ FlatRev s_rev[kFlatWarps][32];
    __shared__ int s_start[kFlatWarps][32];
    __shared__ unsigned char s_nzl[kFlatWarps][32];
    __shared__ unsigned s_sum[kFlatWarps][32][2];  // 24-bit split: native 32-bit shared-memory atomics
    __shared__ float2 s_tab[128];                  // Exp(1) sampler table, staged from global
    if (threadIdx.x < 128) s_tab[threadIdx.x] = kNeglogTab[threadIdx.x];
    __syncthreads();

    const int K = a.kw.K;
    const int64_t total = (int64_t)a.E * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * kFlatWarps + warp;
    const int64_t n_warps = (int64_t)gridDim.x * kFlatWarps;
    // Work items: batches of 32 units.  With scratch.work_counter the warps pull them from an
    // atomic counter (the next pull is requested while the current chunk runs): a warp that was
    // held up -- the last finisher of an env runs the env tail and the drift of its K keywords --
    // simply takes fewer items.  Without the counter: static rounds of 32-unit batches, then the
    // remainder in 8-unit batches dealt round-robin.
    const bool dynamic = a.scratch.work_counter != nullptr && total < (1LL << 34);  // 32-bit pull indices
    uint32_t *const work = a.scratch.work_counter + (a.parity & 1u);
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const unsigned FULL = 0xFFFFFFFFu;
    if (blockIdx.x == 0 && threadIdx.x == 0) reset_next_counters(a);

    FlatCost *costs = s_cost[warp];
    FlatRev *revs = s_rev[warp];
    int *start = s_start[warp];
    unsigned char *nzl = s_nzl[warp];

    auto pull = [&]() -> uint32_t { return lane == 0 ? atomicAdd(work, 1u) : 0u; };
    uint32_t pulled = dynamic ? pull() : 0u;
    int64_t static_pull = gwarp;
    FlatChunk ck;
    ck.u0 = 0; ck.cnt = 0; ck.left = 0;
    for (;;) {
        if (ck.left == 0) {
            int64_t pi = static_pull;
            static_pull += n_warps;
            if (dynamic) {
                pi = (int64_t)__shfl_sync(FULL, pulled, 0);
                pulled = pull();
            }
            ck = flat_chunk(total, n_warps, dynamic, pi);
            if (ck.left == 0) break;
        }
        --ck.left;
        // ---------------- per-unit setup, lane <-> unit ----------------
        const int cnt = ck.cnt;
        const int64_t u = ck.u0 + lane;
        ck.u0 += cnt;
        bool valid = lane < cnt && u < total;
        if (kFloor && a.scratch.outbid_mask != nullptr) {
            // shared auctions: the pre-pass (adc_outbid_rows_kernel) finished the units that cannot win
            if (valid && a.scratch.outbid_mask[u] != 0) valid = false;
            if (!__any_sync(FULL, valid)) continue;
        }
        int e = 0, k = 0, V = 0;
        bool over_cap = false;
        uint32_t genv = 0;
        Unit2 u2;
        u2.t1 = u2.t2 = u2.t3 = u2.full = u2.h1 = u2.a1 = u2.a2 = 0u; u2.L = 0.f; u2.b = 0.f; u2.W = 1;
        float rev_mean = 0.f, rev_sd = 0.f;
        int floor_c = 0;
        if (valid) {
            e = (int)(u / K);
            k = (int)(u - (int64_t)e * K);
            genv = philox_env(a, e);
            const int bid_c = bid_to_cents(load_f(a.bids, a.bids_dtype, u));
            const int win_c = win_cents_of(a, bid_c);
            // an env whose budget bound in its previous step goes straight to the exact walk (which
            // evaluates the same function): nothing of it is computed here
            const bool hinted = a.scratch.serial_hint != nullptr && a.scratch.serial_hint[e] != 0;
            bool outbid = false;  // shared auctions: a rival bids at least as much, no auction can be won
            if (kFloor && !hinted) {
                floor_c = unit_floor(a, e, k, bid_c);
                outbid = bid_c <= floor_c;  // (7 of 8 bidder rows of an 8-bidder world stop here)
                floor_c = max(floor_c, 0);
            }
            if (hinted) {
                over_cap = true;
            } else if (!outbid) {
                const int64_t pi = (int64_t)e * a.kw.env_stride + k;
                const uint4 w = philox4x32_10(0u, a.step, stream_word(ST_UNIT, 0u, (uint32_t)k), genv, k0, k1);
                const long long v = volume_draw(w.x, a.kw.vol_mean[pi], a.kw.vol_std[pi]);
                over_cap = v > kMaxFlatVolume || win_c > kMaxFlatBidCents || v * win_c > kMaxFlatSpendCents;
                V = over_cap ? 0 : (int)v;
                u2 = unit2_make(a.kw.p1[pi], a.kw.p2[pi], a.kw.ctr[pi], a.kw.cvr[pi], win_c);
                rev_mean = (float)a.kw.rev_mean[pi];
                rev_sd = (float)a.kw.rev_std[pi];
            }
        }
        int I = 0, B = 0, S = 0;
        long long cost = 0, rev = 0;
        // shared auctions: a batch in which every unit is outbid (7 of the 8 bidder rows of an 8-bidder world)
        // has no auction to evaluate and goes straight to its outputs
        if (!kFloor || __any_sync(FULL, V > 0)) {
        // ---------------- outcomes: groups of 32 auctions ----------------
        {
            const PhiloxPre pa = philox_pre(a.step, stream_word(ST_AUCTION, 0u, (uint32_t)k), genv, k0, k1);
            const int G = (V + 31) >> 5;
            const int Gmax = __reduce_max_sync(FULL, G);
            const int Gsum = kSpread ? (int)__reduce_add_sync(FULL, (unsigned)G) : 0;
            if (kSpread && ((Gsum + 31) >> 5) + 1 < Gmax) {
                // Uneven days (dense keywords: volumes 128 +- 60 in one batch): lane <-> unit would run the
                // longest day's group count for everybody.  When that costs at least two trips more, the
                // batch's (unit, group) pairs are spread over the lanes instead, like the price draws below:
                // same Philox calls, same masks.
                FlatGroups fg;
                fg.t1 = u2.t1; fg.t2 = u2.t2; fg.t3 = u2.t3; fg.full = u2.full;
                fg.n0 = pa.n0; fg.n1 = pa.n1; fg.x3 = pa.x3; fg.V = V;
                // (the units' group parameters live where their revenue parameters will: the revenue phase
                // writes its own after the prices)
                static_assert(sizeof(FlatGroups) == sizeof(FlatRev), "FlatGroups overlays FlatRev");
                FlatGroups *const grps = reinterpret_cast<FlatGroups *>(revs);
                grps[lane] = fg;
                s_sum[warp][lane][0] = 0u; s_sum[warp][lane][1] = 0u;
                const FlatMap gm = flat_map_begin(G, lane, start, nzl);
                for (int base = 0; base < gm.total; base += 32) {
                    const int b = flat_map_unit(gm, base, lane, nzl);
                    const int i = base + lane;
                    if (i < gm.total) {
                        const FlatGroups f = grps[b];
                        const int g = i - start[b];
                        const int rem = f.V - 32 * g;
                        const uint32_t active = rem >= 32 ? 0xFFFFFFFFu : (1u << rem) - 1u;  // rem >= 1
                        const Masks3 m = group_masks(active, (uint32_t)g, f.t1, f.t2, f.t3, f.full, f.n0, f.n1, f.x3, k0, k1);
                        atomicAdd(&s_sum[warp][b][0], (unsigned)__popc(m.win));
                        atomicAdd(&s_sum[warp][b][1], (unsigned)__popc(m.click) | ((unsigned)__popc(m.conv) << 16));  // <= 65535 each
                    }
                }
                __syncwarp();
                I = (int)s_sum[warp][lane][0];
                B = (int)(s_sum[warp][lane][1] & 0xFFFFu);
                S = (int)(s_sum[warp][lane][1] >> 16);
                __syncwarp();
            } else {
                for (int g = 0; g < Gmax; ++g) {  // lane <-> unit
                    const int rem = V - 32 * g;  // auctions of this group that exist
                    const uint32_t active = rem >= 32 ? 0xFFFFFFFFu : (rem > 0 ? (1u << rem) - 1u : 0u);
                    const Masks3 m = group_masks(active, (uint32_t)g, u2.t1, u2.t2, u2.t3, u2.full, pa.n0, pa.n1, pa.x3, k0, k1);
                    I += __popc(m.win);
                    B += __popc(m.click);
                    S += __popc(m.conv);
                }
            }
        }
        // ---------------- prices: one per click, 4 per Philox call, flattened over the batch ----------------
        {
            const PhiloxPre pc = philox_pre(a.step, stream_word(ST_COST, 0u, (uint32_t)k), genv, k0, k1);
            FlatCost fc;
            fc.t1 = u2.t1; fc.h1 = u2.h1; fc.a1 = u2.a1; fc.a2 = u2.a2; fc.L = u2.L; fc.b = u2.b;
            fc.W = u2.W | ((u2.full & 1u) ? (int)0x80000000u : 0);
            fc.floor_c = floor_c; fc.n0 = pc.n0; fc.n1 = pc.n1; fc.x3 = pc.x3; fc.B = B;
            costs[lane] = fc;
            s_sum[warp][lane][0] = 0u; s_sum[warp][lane][1] = 0u;
        }
        {
            const FlatMap fm = flat_map_begin((B + 3) >> 2, lane, start, nzl);
            const int TB = fm.total;
            for (int base = 0; base < TB; base += 32) {
                const int b = flat_map_unit(fm, base, lane, nzl);
                const int i = base + lane;
                if (i < TB) {
                    const FlatCost fc = costs[b];
                    const int q = i - start[b];
                    const uint4 w = philox_from_pre((uint32_t)q, fc.n0, fc.n1, fc.x3, k0, k1);
                    const bool t1_full = fc.W < 0;
                    const int W = fc.W & 0x7FFFFFFF;
                    const int left = fc.B - 4 * q;  // >= 1 clicks priced by this call
                    int c0 = cost_cents2(w.x, fc.t1, t1_full, fc.h1, fc.a1, fc.a2, fc.L, fc.b, W, s_tab);
                    int c1 = cost_cents2(w.y, fc.t1, t1_full, fc.h1, fc.a1, fc.a2, fc.L, fc.b, W, s_tab);
                    int c2 = cost_cents2(w.z, fc.t1, t1_full, fc.h1, fc.a1, fc.a2, fc.L, fc.b, W, s_tab);
                    int c3 = cost_cents2(w.w, fc.t1, t1_full, fc.h1, fc.a1, fc.a2, fc.L, fc.b, W, s_tab);
                    if (kFloor) {
                        c0 = max(c0, fc.floor_c); c1 = max(c1, fc.floor_c);
                        c2 = max(c2, fc.floor_c); c3 = max(c3, fc.floor_c);
                    }
                    const long long sum = (long long)c0 + (left > 1 ? c1 : 0) + (long long)(left > 2 ? c2 : 0) +
                                          (left > 3 ? c3 : 0);
                    atomicAdd(&s_sum[warp][b][0], (unsigned)(sum & 0xFFFFFF));
                    atomicAdd(&s_sum[warp][b][1], (unsigned)(sum >> 24));
                }
            }
            __syncwarp();
            cost = (long long)s_sum[warp][lane][0] + ((long long)s_sum[warp][lane][1] << 24);
        }
        // ---------------- revenues: one draw per conversion, 4 per Philox call ----------------
        {
            const PhiloxPre pr = philox_pre(a.step, stream_word(ST_REVENUE, 0u, (uint32_t)k), genv, k0, k1);
            FlatRev fr;
            fr.mean = rev_mean; fr.sd = rev_sd; fr.S = S;
            fr.n0 = pr.n0; fr.n1 = pr.n1; fr.x3 = pr.x3; fr.pad0 = 0; fr.pad1 = 0;
            revs[lane] = fr;
            __syncwarp();
            s_sum[warp][lane][0] = 0u; s_sum[warp][lane][1] = 0u;
        }
        const FlatMap rm = flat_map_begin((S + 3) >> 2, lane, start, nzl);
        const int TR = rm.total;
        for (int base = 0; base < TR; base += 32) {
            const int b = flat_map_unit(rm, base, lane, nzl);
            const int i = base + lane;
            if (i < TR) {
                const FlatRev fr = revs[b];
                const int blk = i - start[b];
                const uint4 w = philox_from_pre((uint32_t)blk, fr.n0, fr.n1, fr.x3, k0, k1);
                // all four draws of the call are evaluated (their table loads overlap); the ones
                // past the unit's last conversion are dropped
                const int left = fr.S - 4 * blk;  // >= 1
                const int c0 = revenue_cents(w.x, fr.mean, fr.sd), c1 = revenue_cents(w.y, fr.mean, fr.sd);
                const int c2 = revenue_cents(w.z, fr.mean, fr.sd), c3 = revenue_cents(w.w, fr.mean, fr.sd);
                const long long sum = (long long)c0 + (left > 1 ? c1 : 0) + (long long)(left > 2 ? c2 : 0) +
                                      (left > 3 ? c3 : 0);
                atomicAdd(&s_sum[warp][b][0], (unsigned)(sum & 0xFFFFFF));
                atomicAdd(&s_sum[warp][b][1], (unsigned)(sum >> 24));
            }
        }
        __syncwarp();
        rev = (long long)s_sum[warp][lane][0] + ((long long)s_sum[warp][lane][1] << 24);
        }

        // ---------------- outputs (coalesced: 32 consecutive units), env completion ----------------
        int safe = 0;
        if (valid) {
            a.out.impressions[u] = I;
            a.out.clicks[u] = B;
            a.out.conversions[u] = S;
            a.out.cost_cents[u] = cost;
            a.out.revenue_cents[u] = rev;
            store_f(a.out.cost, a.out.float_dtype, u, cents_to_dollars(cost));
            store_f(a.out.revenue, a.out.float_dtype, u, cents_to_dollars(rev));
            store_flat_unit(a, e, k, I, B, S, cents_to_dollars(cost), cents_to_dollars(rev));
            if (a.out.rows != nullptr) pack_row_unit(a, e, k, I, B, S, cents_to_dollars(cost), cents_to_dollars(rev));
            if (a.out.unit_records != nullptr) store_unit_record(a, u, I, B, S, cents_to_dollars(cost), cents_to_dollars(rev));
            safe = unit_done(a, e, rev - cost, cost, over_cap);
            if (safe && a.out.rows != nullptr) pack_row_tail(a, e);  // this lane ran the env tail
        }
        if (a.drift.mask != nullptr || a.out.episode_profit_cents != nullptr) {
            // finished budget-safe envs: episode accumulation and drift (env:246), the warp shares each
            // env's keywords
            unsigned todo = __ballot_sync(FULL, safe != 0);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int ee = __shfl_sync(FULL, e, src);
                episode_accumulate(a, ee, lane, 32);
                if (a.drift.mask == nullptr) continue;
                const uint32_t ge = philox_env(a, ee);
                // two keywords per lane and trip, all of their loads in flight before the first
                // store: the env's K keywords cost K / 64 DRAM round trips
                for (int kk = lane; kk < K; kk += 64) {
                    const int kb = kk + 32;
                    const bool wa = drift_wanted(a, kk), wb = kb < K && drift_wanted(a, kb);
                    DriftState sa = {0.0, 0.0, 0.0, 0.0}, sb = {0.0, 0.0, 0.0, 0.0};
                    if (wa) sa = drift_load(a, ee, kk);
                    if (wb) sb = drift_load(a, ee, kb);
                    if (wa) {
                        const uint4 w = philox4x32_10(0u, a.step, stream_word(ST_UNIT, 0u, (uint32_t)kk), ge, k0, k1);
                        drift_store(a, ee, kk, sa, drift_from_words(a, w));
                    }
                    if (wb) {
                        const uint4 w = philox4x32_10(0u, a.step, stream_word(ST_UNIT, 0u, (uint32_t)kb), ge, k0, k1);
                        drift_store(a, ee, kb, sb, drift_from_words(a, w));
                    }
                }
            }
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------
// explicit keywords, free-running, budget cannot bind: the flattened step.
//
// An ExplicitKeyword's day (classes:493-538) is one Philox call per auction (impression / click /
// conversion coin flips and the click's price) plus one per sub-step that had no impression (the phantom
// zero-cost slot, classes:514-515) -- with the default env's ~16 auctions a day that is 16 + 23 calls
// per unit, all independent.  A warp takes a batch of `cnt` units (32 on large steps, fewer on small
// ones so that every SM has warps: the default env is 4096 x 10 units) and spreads
//   the auctions of the batch over its lanes (item i belongs to the unit whose prefix range holds i):
//     counts by shared-memory atomics, the sub-steps that saw an impression as a 24-bit mask, and the
//     unit's un-rounded f64 cost sum IN AUCTION ORDER (env:235 sums the day's concatenated clicks): the
//     first lane of every unit segment of the trip adds the segment's clicked prices one after the other;
//   the phantom slots of the batch (the i-th empty sub-step of a unit: fns on the mask);
//   one revenue per conversion, 4 per Philox call, as in the implicit kernel.
// Same draws, same results as one thread walking the unit (adc_units_kernel / the C oracle).
// ------------------------------------------------------------------------------------------
struct __align__(16) FlatExp {  // 64 B per unit in shared memory
    uint32_t thr_impr, thr_click, thr_conv, genv;
    int k, n0, q, pad0;
    double mean, sd;  // price = clamp(mean + sd * z, 0, 4.4) (explicit_cost)
    uint32_t pad1[4];
};
static_assert(sizeof(FlatExp) == 64, "FlatExp layout");

constexpr int kMaxFlatExplicitVolume = 1 << 16;  // beyond: the env takes the exact serial kernel

__global__ void __launch_bounds__(kFlatWarps * 32)
adc_flat_explicit_kernel(const __grid_constant__ adc_step_args a, int cnt)
{
    __shared__ FlatExp s_exp[kFlatWarps][32];
    __shared__ FlatRev s_rev[kFlatWarps][32];
    __shared__ int s_start[kFlatWarps][32];
    __shared__ unsigned char s_nzl[kFlatWarps][32];
    __shared__ int s_cnt[kFlatWarps][3][32];       // impressions, clicks, conversions
    __shared__ unsigned s_mask[kFlatWarps][32];    // sub-steps with an impression; then: sub-steps that get a phantom slot
    __shared__ double s_acc[kFlatWarps][32];       // the units' running f64 cost sums
    __shared__ double s_trip[kFlatWarps][32];      // the prices of the current trip
    __shared__ unsigned s_sum[kFlatWarps][32][2];  // revenue sums, 24-bit split

    const int K = a.kw.K;
    const int64_t total = (int64_t)a.E * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * kFlatWarps + warp;
    const int64_t n_warps = (int64_t)gridDim.x * kFlatWarps;
    const int64_t n_batches = (total + cnt - 1) / cnt;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const unsigned FULL = 0xFFFFFFFFu;
    const adc_tape *no_tape = nullptr;
    if (blockIdx.x == 0 && threadIdx.x == 0) reset_next_counters(a);

    FlatExp *exps = s_exp[warp];
    int *start = s_start[warp];
    unsigned char *nzl = s_nzl[warp];

    for (int64_t batch = gwarp; batch < n_batches; batch += n_warps) {
        // ---------------- per-unit setup, lane <-> unit ----------------
        const int64_t u = batch * cnt + lane;
        const bool valid = lane < cnt && u < total;
        int e = 0, k = 0, V = 0;
        bool force = false;
        float rev_mean = 0.f, rev_sd = 0.f;
        FlatExp fx;
        fx.thr_impr = fx.thr_click = fx.thr_conv = fx.genv = 0u; fx.k = 0; fx.n0 = 0; fx.q = 1; fx.pad0 = 0;
        fx.mean = 0.0; fx.sd = 0.0; fx.pad1[0] = fx.pad1[1] = fx.pad1[2] = fx.pad1[3] = 0u;
        if (valid) {
            e = (int)(u / K);
            k = (int)(u - (int64_t)e * K);
            const UnitPar p = load_unit_par(a, e, k);
            PhiloxSrc src{k0, k1, a.step, philox_env(a, e)};
            uint4 uw;
            const long long Vl = unit_volume(a, src, no_tape, e, k, &uw);
            force = Vl > kMaxFlatExplicitVolume;
            V = force ? 0 : (int)Vl;
            fx.thr_impr = p.thr_impr; fx.thr_click = p.thr_click; fx.thr_conv = p.thr_conv; fx.genv = src.env;
            fx.k = k;
            fx.q = V / ADC_SUBSTEPS;
            fx.n0 = V - (ADC_SUBSTEPS - 1) * fx.q;  // bsim:151-167
            const double xs = __dsqrt_rn(p.bid);     // explicit_cost's constants, once per unit
            fx.sd = __dadd_rn(1e-10, __ddiv_rn(xs, 6.0));
            fx.mean = __dadd_rn(__ddiv_rn(xs, 4.0), 4.4 / 2.0);
            rev_mean = p.rev_mean;
            rev_sd = p.rev_sd;
        }
        __syncwarp();
        exps[lane] = fx;
        s_cnt[warp][0][lane] = 0; s_cnt[warp][1][lane] = 0; s_cnt[warp][2][lane] = 0;
        s_mask[warp][lane] = 0u;
        s_acc[warp][lane] = 0.0;
        // ---------------- auctions, flattened over the batch ----------------
        {
            const FlatMap fm = flat_map_begin(V, lane, start, nzl);
            const int TA = fm.total;
            for (int base = 0; base < TA; base += 32) {
                const int b = flat_map_unit(fm, base, lane, nzl);
                const int i = base + lane;
                const bool act = i < TA;
                bool clk = false;
                double cost = 0.0;
                if (act) {
                    const FlatExp f = exps[b];
                    const int j = i - start[b];
                    const uint4 w = philox4x32_10((uint32_t)j, a.step, stream_word(ST_AUCTION, 0u, (uint32_t)f.k), f.genv, k0, k1);
                    if (w.x <= f.thr_impr) {  // Bernoulli(p); the lane's sum is Binomial(n, p)
                        const int t = j < f.n0 ? 0 : 1 + (j - f.n0) / f.q;
                        atomicOr(&s_mask[warp][b], 1u << t);
                        atomicAdd(&s_cnt[warp][0][b], 1);
                        clk = w.y <= f.thr_click;
                        if (clk) {
                            cost = clampd(__dadd_rn(f.mean, __dmul_rn(f.sd, (double)znorm(w.w))), 0.0, 4.4);
                            atomicAdd(&s_cnt[warp][1][b], 1);
                            if (w.z <= f.thr_conv) atomicAdd(&s_cnt[warp][2][b], 1);
                        }
                    }
                }
                // the units' cost sums, in auction order: lanes of one unit are neighbours
                s_trip[warp][lane] = cost;
                __syncwarp();
                const unsigned clicked = __ballot_sync(FULL, clk);
                const int b_prev = __shfl_up_sync(FULL, b, 1);
                const bool first = act && (lane == 0 || b != b_prev);
                const unsigned firsts = __ballot_sync(FULL, first);
                if (first) {
                    const unsigned above = lane == 31 ? 0u : firsts & (0xFFFFFFFFu << (lane + 1));
                    const unsigned upto = above ? (1u << (__ffs(above) - 1)) - 1u : 0xFFFFFFFFu;
                    unsigned seg = clicked & upto & (0xFFFFFFFFu << lane);
                    if (seg) {
                        double sum = s_acc[warp][b];
                        while (seg) {
                            const int l = __ffs(seg) - 1;
                            seg &= seg - 1;
                            sum = __dadd_rn(sum, s_trip[warp][l]);
                        }
                        s_acc[warp][b] = sum;
                    }
                }
                __syncwarp();
            }
        }
        // ---------------- phantom slots: the sub-steps without an impression (classes:514-515) ----------------
        {
            __syncwarp();
            const unsigned pm = valid && !force ? ~s_mask[warp][lane] & ((1u << ADC_SUBSTEPS) - 1u) : 0u;
            __syncwarp();
            s_mask[warp][lane] = pm;
            const FlatMap fm = flat_map_begin(__popc(pm), lane, start, nzl);
            const int TP = fm.total;
            for (int base = 0; base < TP; base += 32) {
                const int b = flat_map_unit(fm, base, lane, nzl);
                const int i = base + lane;
                if (i < TP) {
                    const FlatExp f = exps[b];
                    const int t = (int)__fns(s_mask[warp][b], 0u, i - start[b] + 1);
                    const uint4 w = philox4x32_10((uint32_t)t, a.step, stream_word(ST_PHANTOM, 0u, (uint32_t)f.k), f.genv, k0, k1);
                    if (w.y <= f.thr_click) {
                        atomicAdd(&s_cnt[warp][1][b], 1);
                        if (w.z <= f.thr_conv) atomicAdd(&s_cnt[warp][2][b], 1);
                    }
                }
            }
            __syncwarp();
        }
        const int I = s_cnt[warp][0][lane], B = s_cnt[warp][1][lane], S = s_cnt[warp][2][lane];
        const double cost_f = s_acc[warp][lane];
        // ---------------- revenues: one draw per conversion, 4 per Philox call ----------------
        {
            const PhiloxPre pr = philox_pre(a.step, stream_word(ST_REVENUE, 0u, (uint32_t)k), fx.genv, k0, k1);
            FlatRev fr;
            fr.mean = rev_mean; fr.sd = rev_sd; fr.S = S;
            fr.n0 = pr.n0; fr.n1 = pr.n1; fr.x3 = pr.x3; fr.pad0 = 0; fr.pad1 = 0;
            s_rev[warp][lane] = fr;
            s_sum[warp][lane][0] = 0u; s_sum[warp][lane][1] = 0u;
        }
        const FlatMap rm = flat_map_begin((S + 3) >> 2, lane, start, nzl);
        const int TR = rm.total;
        for (int base = 0; base < TR; base += 32) {
            const int b = flat_map_unit(rm, base, lane, nzl);
            const int i = base + lane;
            if (i < TR) {
                const FlatRev fr = s_rev[warp][b];
                const int blk = i - start[b];
                const uint4 w = philox_from_pre((uint32_t)blk, fr.n0, fr.n1, fr.x3, k0, k1);
                const int left = fr.S - 4 * blk;  // >= 1
                const int c0 = revenue_cents(w.x, fr.mean, fr.sd), c1 = revenue_cents(w.y, fr.mean, fr.sd);
                const int c2 = revenue_cents(w.z, fr.mean, fr.sd), c3 = revenue_cents(w.w, fr.mean, fr.sd);
                const long long sum = (long long)c0 + (left > 1 ? c1 : 0) + (long long)(left > 2 ? c2 : 0) +
                                      (left > 3 ? c3 : 0);
                atomicAdd(&s_sum[warp][b][0], (unsigned)(sum & 0xFFFFFF));
                atomicAdd(&s_sum[warp][b][1], (unsigned)(sum >> 24));
            }
        }
        __syncwarp();
        const long long rev_c = (long long)s_sum[warp][lane][0] + ((long long)s_sum[warp][lane][1] << 24);

        // ---------------- outputs, env completion ----------------
        int safe = 0;
        if (valid) {
            a.out.impressions[u] = I;
            a.out.clicks[u] = B;
            a.out.conversions[u] = S;
            a.out.revenue_cents[u] = rev_c;
            store_f(a.out.revenue, a.out.float_dtype, u, cents_to_dollars(rev_c));
            a.out.cost_cents[u] = 0;
            // a unit beyond the kernel's volume cap makes the env look unaffordable: the serial kernel walks it
            a.scratch.unit_cost_f64[u] = force ? __longlong_as_double(0x7FF0000000000000LL) : cost_f;
            store_f(a.out.cost, a.out.float_dtype, u, cost_f);
            store_flat_unit(a, e, k, I, B, S, cost_f, cents_to_dollars(rev_c));
            safe = unit_done(a, e, 0, 0);
        }
        if (a.drift.mask != nullptr || a.out.episode_profit_cents != nullptr) {
            unsigned todo = __ballot_sync(FULL, safe != 0);
            while (todo) {
                const int src_lane = __ffs(todo) - 1;
                todo &= todo - 1;
                const int ee = __shfl_sync(FULL, e, src_lane);
                episode_accumulate(a, ee, lane, 32);
                if (a.drift.mask == nullptr) continue;
                const uint32_t ge = philox_env(a, ee);
                for (int kk = lane; kk < K; kk += 32) {
                    if (!drift_wanted(a, kk)) continue;
                    const uint4 w = philox4x32_10(0u, a.step, stream_word(ST_UNIT, 0u, (uint32_t)kk), ge, k0, k1);
                    drift_apply(a, ee, kk, drift_from_words(a, w));
                }
            }
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------
// replay kernel (implicit keywords, tape-driven, budget cannot bind): HBM-bound by construction.
// One warp per unit; every tape stream is read with coalesced loads:
//   competitor bids  32 x int32 per trip, 4 trips in flight;
//   click uniforms   gathered by impression rank (ballot + popc), contiguous among the winners;
//   conversion uniforms and revenues are consumed by rank only, so they are dense streaming
//   reductions over [0, clicks) and [0, conversions).
// Algorithmic bytes per unit: 4 V + 8 I + 8 B + 4 S (streams) + 36 (volume, 4 CSR offsets) +
// 4 (bid) + 36 (outputs).
// ------------------------------------------------------------------------------------------
struct __align__(16) ReplayUnit {  // 80 B per unit in shared memory
    const int32_t *comp;   // stream base pointers of the unit
    const double *click;
    const double *conv;
    const int32_t *rev;
    double ctr, cvr;
    int V, bid_cents;
    int n_comp, n_click, n_conv, n_rev;  // stream lengths (bounds of a truncated tape)
};

constexpr int kReplayWarps = 8;

__device__ __forceinline__ int clamp_len(int64_t n) { return n > 0x7FFFFFFF ? 0x7FFFFFFF : (int)n; }

// Warp-batched like the hot kernel: lane <-> unit for the header (coalesced loads of volume,
// CSR offsets, bid, rates) and for the outputs, all 32 lanes on one unit's streams in between.
// The dependent chain of a unit is  competitor bids -> click uniforms (by impression rank) ->
// conversion uniforms (count = clicks) -> revenues (count = conversions); the first 64
// conversion uniforms and revenues are fetched speculatively together with the competitor bids,
// which removes two of the four DRAM round trips for a typical day (~38 clicks, ~31
// conversions) at < 20 % over-read.
__global__ void __launch_bounds__(kReplayWarps * 32)
adc_replay_implicit_kernel(const __grid_constant__ adc_step_args a, const __grid_constant__ adc_tape t)
{
    __shared__ ReplayUnit s_unit[kReplayWarps][32];
    const int K = a.kw.K;
    const int64_t total = (int64_t)a.E * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * kReplayWarps + warp;
    const int64_t n_warps = (int64_t)gridDim.x * kReplayWarps;
    const int64_t n_batches = (total + 31) / 32;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lt = (1u << lane) - 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0) reset_next_counters(a);
    ReplayUnit *units = s_unit[warp];

    for (int64_t batch = gwarp; batch < n_batches; batch += n_warps) {
        // ---------------- header, lane <-> unit ----------------
        const int64_t u = batch * 32 + lane;
        const bool valid = u < total;
        int e = 0, myV = 0;
        {
            ReplayUnit ru;
            ru.comp = t.comp_cents; ru.click = t.u_click; ru.conv = t.u_conv; ru.rev = t.rev_cents;
            ru.ctr = 0.0; ru.cvr = 0.0; ru.V = 0; ru.bid_cents = 0;
            ru.n_comp = ru.n_click = ru.n_conv = ru.n_rev = 0;
            if (valid) {
                e = (int)(u / K);
                const int64_t pi = (int64_t)e * a.kw.env_stride + (u - (int64_t)e * K);
                const int64_t c0 = t.comp_off[u], k0 = t.click_off[u], v0 = t.conv_off[u], r0 = t.rev_off[u];
                ru.comp += c0; ru.click += k0; ru.conv += v0; ru.rev += r0;
                ru.n_comp = clamp_len(t.comp_off[u + 1] - c0);
                ru.n_click = clamp_len(t.click_off[u + 1] - k0);
                ru.n_conv = clamp_len(t.conv_off[u + 1] - v0);
                ru.n_rev = clamp_len(t.rev_off[u + 1] - r0);
                ru.ctr = a.kw.ctr[pi];
                ru.cvr = a.kw.cvr[pi];
                ru.V = t.volume[u];
                ru.bid_cents = win_cents_of(a, bid_to_cents(load_f(a.bids, a.bids_dtype, u)));
                myV = ru.V;
            }
            units[lane] = ru;
        }
        __syncwarp();

        // ---------------- the 32 units, one after the other, 32 lanes per unit ----------------
        // Software-pipelined: the first 128 competitor bids and the speculative heads of the
        // conversion / revenue streams of the NEXT non-empty unit are requested before the current
        // unit is consumed, so two units' worth of loads are in flight per warp.
        int I = 0, B = 0, S = 0;
        long long cost = 0, rev = 0;
        bool my_overrun = false;
        struct Pre {
            int c[4];
            double cv[2];
            int rv[2];
            bool overrun;
        };
        auto issue = [&](int b, Pre &p) {
            const ReplayUnit &h = units[b];
            p.overrun = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = 32 * q + lane;
                p.c[q] = 0x7FFFFFFF;
                if (j < h.V) {
                    if (j < h.n_comp) p.c[q] = __ldg(h.comp + j); else p.overrun = true;
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = lane + 32 * q;
                p.cv[q] = i < h.n_conv ? __ldg(h.conv + i) : 2.0;
                p.rv[q] = i < h.n_rev ? __ldg(h.rev + i) : 0;
            }
        };
        unsigned todo = __ballot_sync(FULL, myV > 0);
        Pre nxt;
        int b_next = todo ? __ffs(todo) - 1 : -1;
        if (b_next >= 0) issue(b_next, nxt);
        while (b_next >= 0) {
            const int b = b_next;
            const Pre cur = nxt;
            todo &= todo - 1;
            b_next = todo ? __ffs(todo) - 1 : -1;
            if (b_next >= 0) issue(b_next, nxt);
            const ReplayUnit h = units[b];
            const int Vb = h.V;
            bool overrun = cur.overrun;
            int nI = 0, Bl = 0;
            long long costw = 0;
            for (int base = 0; base < Vb; base += 128) {
                int c[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (base == 0) {
                        c[q] = cur.c[q];
                    } else {
                        const int j = base + 32 * q + lane;
                        c[q] = 0x7FFFFFFF;
                        if (j < Vb) {
                            if (j < h.n_comp) c[q] = __ldg(h.comp + j); else overrun = true;
                        }
                    }
                }
                int rank[4];
                bool win[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    win[q] = h.bid_cents > c[q];
                    const unsigned m = __ballot_sync(FULL, win[q]);
                    rank[q] = nI + __popc(m & lt);
                    nI += __popc(m);
                }
                double uc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uc[q] = 2.0;
                    if (win[q]) {
                        if (rank[q] < h.n_click) uc[q] = __ldg(h.click + rank[q]); else overrun = true;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool clk = win[q] && uc[q] <= h.ctr;
                    Bl += clk;
                    costw += clk ? c[q] : 0;
                }
            }
            const int nB = (int)__reduce_add_sync(FULL, (unsigned)Bl);
            int Sl = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = lane + 32 * q;
                if (i < nB) {
                    if (i < h.n_conv) Sl += cur.cv[q] <= h.cvr; else overrun = true;
                }
            }
            for (int i = 64 + lane; i < nB; i += 32) {
                if (i < h.n_conv) Sl += __ldg(h.conv + i) <= h.cvr; else overrun = true;
            }
            const int nS = (int)__reduce_add_sync(FULL, (unsigned)Sl);
            long long revl = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = lane + 32 * q;
                if (i < nS) {
                    if (i < h.n_rev) revl += cur.rv[q]; else overrun = true;
                }
            }
            for (int i = 64 + lane; i < nS; i += 32) {
                if (i < h.n_rev) revl += __ldg(h.rev + i); else overrun = true;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                costw += __shfl_xor_sync(FULL, costw, off);
                revl += __shfl_xor_sync(FULL, revl, off);
            }
            overrun = __any_sync(FULL, overrun);
            if (lane == b) {
                I = nI; B = nB; S = nS; cost = costw; rev = revl; my_overrun = overrun;
            }
        }

        // ---------------- outputs (coalesced), env completion ----------------
        int safe = 0;
        if (valid) {
            a.out.impressions[u] = I;
            a.out.clicks[u] = B;
            a.out.conversions[u] = S;
            a.out.cost_cents[u] = cost;
            a.out.revenue_cents[u] = rev;
            store_f(a.out.cost, a.out.float_dtype, u, cents_to_dollars(cost));
            store_f(a.out.revenue, a.out.float_dtype, u, cents_to_dollars(rev));
            store_flat_unit(a, e, (int)(u - (int64_t)e * K), I, B, S, cents_to_dollars(cost), cents_to_dollars(rev));
            safe = unit_done(a, e, rev - cost, cost, my_overrun);
        }
        if (a.drift.mask != nullptr || a.out.episode_profit_cents != nullptr) {
            unsigned todo = __ballot_sync(FULL, safe != 0);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int ee = __shfl_sync(FULL, e, src);
                episode_accumulate(a, ee, lane, 32);
                if (a.drift.mask == nullptr) continue;
                for (int kk = lane; kk < K; kk += 32) {
                    if (!drift_wanted(a, kk)) continue;
                    drift_apply(a, ee, kk, unit_drift<TapeSrc>(a, &t, ee, kk, make_uint4(0, 0, 0, 0)));
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// packed replay kernel (implicit keywords, adc_tape.packed): the HBM-bound path.
//
// A unit's whole day is one 16-byte aligned record (header, competitor bids, click uniforms,
// conversion uniforms, revenues; see include/adcraft_b200.h), so the kernel moves it with ONE bulk
// copy (TMA, cp.async.bulk -> mbarrier complete_tx) into a per-warp byte ring in shared memory, as
// many units ahead of the walk as fit (records are variable-sized: a dense day is ~1.8 KB, so a
// 6 KB ring keeps 3 in flight).  No per-lane address arithmetic, bounds checks or
// dependent DRAM round trips are left in the walk: the warp reads the record from shared memory
// (128-bit loads of four consecutive competitor bids per lane, click uniforms gathered by
// impression rank, conversion uniforms and revenues as dense prefixes) and reduces with REDUX.
// Records that do not fit the ring (very large volumes), bids above kMaxFlatBidCents or values
// outside the 16-bit fast-path range are walked by pk_walk_generic (64-bit sums, any address
// space); malformed records flag an overrun, which routes the env to the serial kernel (CSR tape).
// ------------------------------------------------------------------------------------------
constexpr int kPkMaxCap = 7936;  // largest ring that may be instantiated: bounds the 32-bit fast-path sums (see pk_walk_fast)

struct __align__(16) PkUnit {  // 48 B per unit in shared memory
    const unsigned char *src;  // the record in global memory
    int bytes;                 // record size, 0 = nothing to walk
    int bid_cents;
    double ctr, cvr;
    int n_comp, n_click, n_conv, n_rev;  // validated header of a record that takes the fast walk
};

struct __align__(16) PkOut {  // a walked unit's sums, written over the first 32 B of its PkUnit
    int I, B, S, overrun;
    long long cost, rev;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// The tape streams through once per step: evict-first keeps it from displacing the outputs and the
// per-env accumulators in L2.
// One lane of the (converged) warp, chosen by the hardware: unlike `lane == 0` the compiler knows that a
// single lane is active behind it and moves a bulk copy's operands to uniform registers without a loop.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}

__device__ __forceinline__ unsigned long long evict_first_policy()
{
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, unsigned long long policy)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    // try_wait suspends the thread for a hardware time slice before it reports failure, so this
    // loop does not burn issue slots; there is no spin bound (a bound that traps kills the context
    // when a profiler replay stretches a copy's latency)
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}

struct PkResult {
    int I, B, S;
    long long cost, rev;
    bool overrun;
};

// Record header -> stream positions; false (and empty streams) when the header does not fit the
// record.  All values are warp-uniform.
struct PkView {
    int V, n_comp, n_click, n_conv, n_rev;
    const unsigned char *comp, *click, *conv, *rev;
};

__device__ __forceinline__ bool pk_view(const unsigned char *rec, int bytes, PkView &v)
{
    const int4 h0 = *reinterpret_cast<const int4 *>(rec);
    v.V = h0.x; v.n_comp = h0.y; v.n_click = h0.z; v.n_conv = h0.w;
    v.n_rev = *reinterpret_cast<const int *>(rec + 16);
    const unsigned lim = (unsigned)bytes;
    bool ok = (unsigned)v.n_comp <= lim && (unsigned)v.n_click <= lim && (unsigned)v.n_conv <= lim &&
              (unsigned)v.n_rev <= lim && v.V >= 0 && v.n_comp <= v.V;
    const long long comp_b = 4LL * ((v.n_comp + 3) & ~3);
    const long long need = 32LL + comp_b + 8LL * v.n_click + 8LL * v.n_conv + 4LL * v.n_rev;
    ok = ok && need <= (long long)bytes;
    if (!ok) { v.n_comp = v.n_click = v.n_conv = v.n_rev = 0; }
    v.comp = rec + 32;
    v.click = v.comp + (ok ? comp_b : 0);
    v.conv = v.click + 8LL * v.n_click;
    v.rev = v.conv + 8LL * v.n_conv;
    return ok;
}

// Fast-walk precondition, checked once per unit by its owner lane on the header in global memory:
// every count below 2048 (keeps the 32-bit size arithmetic exact and the 32-bit sums safe), the
// streams fit the record, and the competitor stream covers the whole volume.
__device__ __forceinline__ bool pk_header_ok(const int4 h0, int n_rev, int bytes)
{
    const int V = h0.x, n_comp = h0.y, n_click = h0.z, n_conv = h0.w;
    return (unsigned)(n_comp | n_click | n_conv | n_rev) < 2048u &&
           32 + ((n_comp + 3) & ~3) * 4 + 8 * (n_click + n_conv) + 4 * n_rev <= bytes && V == n_comp;
}

// Any record, any address space, 64-bit sums.
__device__ __noinline__ PkResult pk_walk_generic(const unsigned char *rec, int bytes, int bid_cents, double ctr,
                                                 double cvr, int lane)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lt = (1u << lane) - 1u;
    PkView v;
    PkResult r;
    r.overrun = !pk_view(rec, bytes, v);
    r.overrun = r.overrun || v.V > v.n_comp;
    const int *comp = reinterpret_cast<const int *>(v.comp);
    const double *click = reinterpret_cast<const double *>(v.click);
    const double *conv = reinterpret_cast<const double *>(v.conv);
    const int *rev = reinterpret_cast<const int *>(v.rev);
    int nI = 0, Bl = 0;
    long long costl = 0;
    for (int base = 0; base < v.n_comp; base += 32) {
        const int j = base + lane;
        const int c = j < v.n_comp ? comp[j] : 0x7FFFFFFF;
        const bool win = bid_cents > c;
        const unsigned m = __ballot_sync(FULL, win);
        const int rank = nI + __popc(m & lt);
        nI += __popc(m);
        const bool clk = win && rank < v.n_click && click[rank] <= ctr;
        Bl += clk;
        costl += clk ? c : 0;
    }
    r.overrun = r.overrun || nI > v.n_click;
    r.I = nI;
    r.B = (int)__reduce_add_sync(FULL, (unsigned)Bl);
    r.overrun = r.overrun || r.B > v.n_conv;
    int Sl = 0;
    for (int i = lane; i < min(r.B, v.n_conv); i += 32) Sl += conv[i] <= cvr;
    r.S = (int)__reduce_add_sync(FULL, (unsigned)Sl);
    r.overrun = r.overrun || r.S > v.n_rev;
    long long revl = 0;
    for (int i = lane; i < min(r.S, v.n_rev); i += 32) revl += rev[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        costl += __shfl_xor_sync(FULL, costl, off);
        revl += __shfl_xor_sync(FULL, revl, off);
    }
    r.cost = costl;
    r.rev = revl;
    return r;
}

// Shared-memory loads by 32-bit shared-space address.  (Through generic pointers the compiler
// re-derives the CTA's shared window base -- SR_CgaCtaId, two adds, a LEA -- next to every predicated
// load of the walk: a fifth of its instructions.)
__device__ __forceinline__ int4 lds_v4(uint32_t a)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a)
{
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}

// Record resident in shared memory (`rec`: its shared-space address), bid <= kMaxFlatBidCents, header flag
// ADC_PACKED_NARROW (no negative competitor bid, revenues below 65536 cents: checked by the caller): 32-bit
// address and sum arithmetic, REDUX reductions.  Returns false when the record needs pk_walk_generic
// instead: click uniforms running out mid-walk.  The 32-bit sums are exact: a record of at most kPkMaxCap
// bytes holds at most 1976 bids = 15 full trips and 2 single ones, so a lane adds at most 62 16-bit values below
// its click counter at bit 22.  No load is conditional: a lane without a bid or a won auction
// reads whatever follows (at most 544 bytes past the record's end: the launch pads the rings for it) and
// ignores it.
__device__ __forceinline__ bool pk_walk_fast(uint32_t rec, int n_comp, int n_click, int n_conv, int n_rev,
                                             int bid_cents, double ctr, double cvr, int lane, unsigned lt, PkResult &r)
{
    const unsigned FULL = 0xFFFFFFFFu;
    // the header was validated by the unit's owner lane (pk_header_ok): all counts < 2048 and
    // consistent with the record size, n_comp == V
    r.overrun = false;
    const int n4 = (n_comp + 3) >> 2;  // int4 groups; padding entries are INT32_MAX and never win
    const uint32_t click = rec + 32u + 16u * (uint32_t)n4;
    const uint32_t conv = click + 8u * (uint32_t)n_click, rev = conv + 8u * (uint32_t)n_conv;
    constexpr unsigned kOne = 1u << 22;  // click counter above the lane's cost sum (< 64 * 65536)
    int nI = 0;
    unsigned acc = 0;
    // full trips: 128 auctions, four consecutive ones per lane (one 128-bit load)
    uint32_t cp = rec + 32u + 16u * (uint32_t)lane;
    int left = n4 - lane;  // > 0: the lane has bids in the trip
    int g0 = 0;
    for (; n4 - g0 > 16; g0 += 32) {
        const bool in = left > 0;
        const int4 c = lds_v4(cp);  // (a lane past the last bid reads at most 512 bytes past the record: padded for)
        const bool w0 = in && bid_cents > c.x, w1 = in && bid_cents > c.y;
        const bool w2 = in && bid_cents > c.z, w3 = in && bid_cents > c.w;
        const unsigned m0 = __ballot_sync(FULL, w0), m1 = __ballot_sync(FULL, w1);
        const unsigned m2 = __ballot_sync(FULL, w2), m3 = __ballot_sync(FULL, w3);
        // auction order inside the trip is lane-major (j = 4 g + q): rank = wins of lower lanes + own earlier wins;
        // the click uniforms of a lane's wins are consecutive
        // (a tape shorter than the walk -- never on a consistent recording -- is noticed after the loops: the
        // index is clamped so that the loads stay within 32 bytes of the record)
        const uint32_t u0 = click + 8u * (uint32_t)min(nI + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt), n_click);
        const uint32_t u1 = u0 + (w0 ? 8u : 0u), u2 = u1 + (w1 ? 8u : 0u), u3 = u2 + (w2 ? 8u : 0u);
        nI += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
        const bool k0 = w0 & (lds_f64(u0) <= ctr), k1 = w1 & (lds_f64(u1) <= ctr);
        const bool k2 = w2 & (lds_f64(u2) <= ctr), k3 = w3 & (lds_f64(u3) <= ctr);
        acc += (k0 ? (unsigned)c.x + kOne : 0u) + (k1 ? (unsigned)c.y + kOne : 0u);
        acc += (k2 ? (unsigned)c.z + kOne : 0u) + (k3 ? (unsigned)c.w + kOne : 0u);
        cp += 512u;
        left -= 32;
    }
    // the last <= 64 auctions: one per lane and trip (a volume just above 128 would otherwise pay a
    // whole 128-wide trip for a handful of auctions)
    for (int j0 = 4 * g0; j0 < n_comp; j0 += 32) {
        const int j = j0 + lane;
        const bool in = j < n_comp;
        const int c1 = lds_s32(rec + 32u + 4u * (uint32_t)j);
        const bool w = in && bid_cents > c1;
        const unsigned m = __ballot_sync(FULL, w);
        const int rk = min(nI + __popc(m & lt), n_click);
        nI += __popc(m);
        const bool k = w & (lds_f64(click + 8u * (uint32_t)rk) <= ctr);
        acc += k ? (unsigned)c1 + kOne : 0u;
    }
    if (nI > n_click) return false;  // click uniforms ran out: pk_walk_generic decides
    r.I = nI;
    r.B = (int)__reduce_add_sync(FULL, acc >> 22);
    r.overrun = r.overrun || r.B > n_conv;
    const int nB = min(r.B, n_conv);
    // (two per lane without a loop: a day of more than 64 accepted clicks is the exception)
    unsigned Sl = (unsigned)((lane < nB) & (lds_f64(lane < nB ? conv + 8u * (uint32_t)lane : rec) <= cvr)) +
                  (unsigned)((lane + 32 < nB) & (lds_f64(lane + 32 < nB ? conv + 8u * (uint32_t)(lane + 32) : rec) <= cvr));
    if (nB > 64) {
#pragma unroll 1
        for (int i = lane + 64; i < nB; i += 32) Sl += lds_f64(conv + 8u * (uint32_t)i) <= cvr;
    }
    r.S = (int)__reduce_add_sync(FULL, Sl);
    r.overrun = r.overrun || r.S > n_rev;
    const int nS = min(r.S, n_rev);
    unsigned revl = (lane < nS ? (unsigned)lds_s32(rev + 4u * (uint32_t)lane) : 0u) +
                    (lane + 32 < nS ? (unsigned)lds_s32(rev + 4u * (uint32_t)(lane + 32)) : 0u);
    if (nS > 64) {
#pragma unroll 1
        for (int i = lane + 64; i < nS; i += 32) revl += (unsigned)lds_s32(rev + 4u * (uint32_t)i);
    }
    r.cost = (long long)__reduce_add_sync(FULL, acc & (kOne - 1u));
    r.rev = (long long)__reduce_add_sync(FULL, revl);
    return true;
}

template <int kPkWarps, int kRing, int kDepth, int kMinBlocks>
__global__ void __launch_bounds__(kPkWarps * 32, kMinBlocks)
adc_replay_packed_kernel(const __grid_constant__ adc_step_args a, const __grid_constant__ adc_tape t)
{
    static_assert(kRing <= kPkMaxCap && kRing % 16 == 0, "ring size");
    static_assert((kDepth & (kDepth - 1)) == 0, "copies in flight: a power of two");
    extern __shared__ __align__(128) unsigned char pk_buf[];  // [kPkWarps][kRing]
    __shared__ PkUnit s_unit[kPkWarps][32];
    __shared__ __align__(8) unsigned long long s_bar[kPkWarps][kDepth];
    const int K = a.kw.K;
    const int64_t total = (int64_t)a.E * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gwarp = (int64_t)blockIdx.x * kPkWarps + warp;
    const int64_t n_warps = (int64_t)gridDim.x * kPkWarps;
    // every warp walks one contiguous range of units (equal shares, +-1 unit) in batches of 32
    const int64_t u_begin = (total / n_warps) * gwarp + min(gwarp, total % n_warps);
    const int64_t u_end = u_begin + total / n_warps + (gwarp < total % n_warps ? 1 : 0);
    const unsigned FULL = 0xFFFFFFFFu;
    if (blockIdx.x == 0 && threadIdx.x == 0) reset_next_counters(a);

    PkUnit *units = s_unit[warp];
    unsigned char *ring = pk_buf + (size_t)warp * kRing;
    const uint32_t ring_s = smem_addr(ring);
    const uint32_t bar_s = smem_addr(&s_bar[warp][0]);
    if (lane < kDepth) mbar_init(bar_s + 8u * lane, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    // Ring state, warp-uniform.  Records are placed in issue order at `head`, wrapping to 0 when the
    // next one does not fit before the end; the consumer replays the same placement rule from
    // `chead`, so no per-record bookkeeping is stored.  [chead, head) in ring order is in flight.
    unsigned head = 0, chead = 0, used = 0, n_iss = 0, n_con = 0;  // (copies issued / consumed so far)
    int in_flight = 0;
    const unsigned long long policy = evict_first_policy();
    const unsigned lt = (1u << lane) - 1u;

    for (int64_t u0 = u_begin; u0 < u_end; u0 += 32) {
        // ---------------- header, lane <-> unit ----------------
        const int64_t u = u0 + lane;
        const bool valid = u < u_end;
        int e = 0;
        PkUnit pu;
        pu.src = t.packed; pu.bytes = 0; pu.bid_cents = 0; pu.ctr = 0.0; pu.cvr = 0.0;
        pu.n_comp = pu.n_click = pu.n_conv = pu.n_rev = 0;
        bool my_overrun = false, fast = false;
        if (valid) {
            e = (int)(u / K);
            const int64_t pi = (int64_t)e * a.kw.env_stride + (u - (int64_t)e * K);
            const long long o0 = t.packed_off[u], o1 = t.packed_off[u + 1];
            const long long len = o1 - o0;
            const bool sane = o0 >= 0 && len >= 0 && len <= 0x7FFFFFF0LL && ((o0 | len) & 15) == 0 && (len == 0 || len >= 32);
            my_overrun = !sane;
            pu.src = t.packed + o0;
            pu.bytes = sane ? (int)len : 0;
            pu.ctr = a.kw.ctr[pi];
            pu.cvr = a.kw.cvr[pi];
            pu.bid_cents = win_cents_of(a, bid_to_cents(load_f(a.bids, a.bids_dtype, u)));
            if (pu.bytes > 0 && pu.bytes <= kRing && pu.bid_cents <= kMaxFlatBidCents) {
                // owner lane reads the record header (the bulk copy re-reads the line from L2)
                const int4 h0 = __ldg(reinterpret_cast<const int4 *>(pu.src));
                const int2 h1 = __ldg(reinterpret_cast<const int2 *>(pu.src + 16));  // n_rev, flags
                const int n_rev = h1.x;
                if (pk_header_ok(h0, n_rev, pu.bytes)) {
                    // (a record without ADC_PACKED_NARROW -- negative competitor bids, revenues beyond 16 bits --
                    // is walked by pk_walk_generic from global memory)
                    fast = (h1.y & ADC_PACKED_NARROW) != 0;
                    pu.n_comp = h0.y; pu.n_click = h0.z; pu.n_conv = h0.w; pu.n_rev = n_rev;
                } else {
                    my_overrun = true;  // the serial kernel re-walks the env from the CSR streams
                    pu.bytes = 0;
                }
            }
        }
        units[lane] = pu;
        __syncwarp();
        const unsigned work_m = __ballot_sync(FULL, pu.bytes > 0);
        const unsigned fast_m = __ballot_sync(FULL, fast);

        // ---------------- the units, one after the other; copies run ahead as far as the ring allows ----------------
        // issue cursor (warp-uniform): the next fast unit to copy and its {src, bytes, bid} words
        unsigned iss_m = fast_m;
        uint4 nx = make_uint4(0, 0, 0, 0);
        if (iss_m) nx = *reinterpret_cast<const uint4 *>(&units[__ffs(iss_m) - 1]);
        auto try_issue = [&]() -> bool {
            const unsigned bytes = nx.z;
            // pointers coincide: restart at the ring base (an empty ring whose head sits behind a record's size
            // could otherwise never take that record)
            if (in_flight == 0) { head = 0; chead = 0; used = 0; }
            // `used` = bytes between chead and head in ring order, skipped ring ends included
            const bool straight = head + bytes <= (unsigned)kRing;
            const unsigned need = straight ? bytes : bytes + ((unsigned)kRing - head);
            if (in_flight >= kDepth || used + need > (unsigned)kRing) return false;
            const unsigned off = straight ? head : 0u;
            if (elect_one()) {
                const unsigned char *src = reinterpret_cast<const unsigned char *>(
                    ((unsigned long long)nx.y << 32) | nx.x);
                bulk_load(ring_s + off, src, bytes, bar_s + 8u * (n_iss & (kDepth - 1)), policy);
            }
            used += need;
            head = off + bytes;
            ++n_iss;
            ++in_flight;
            iss_m &= iss_m - 1;
            if (iss_m) nx = *reinterpret_cast<const uint4 *>(&units[__ffs(iss_m) - 1]);
            return true;
        };
#pragma unroll 1
        while (iss_m && try_issue()) {}
        unsigned todo = work_m;
#pragma unroll 1
        while (todo) {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1;
            __syncwarp();  // everyone is done reading the records consumed so far
            // up to two copies per unit: a copy that did not fit last time (ring end skipped) is caught
            // up here, otherwise the lead over the walk decays to zero within a batch
            if (iss_m && try_issue() && iss_m && in_flight < 3) try_issue();
            const PkUnit h = units[b];
            PkResult r;
            if ((fast_m >> b) & 1u) {
                mbar_wait(bar_s + 8u * (n_con & (kDepth - 1)), (n_con / kDepth) & 1u);  // slot and phase of the n-th copy
                ++n_con;
                const bool straight = chead + (unsigned)h.bytes <= (unsigned)kRing;
                const unsigned off = straight ? chead : 0u;
                used -= straight ? (unsigned)h.bytes : (unsigned)h.bytes + ((unsigned)kRing - chead);
                chead = off + (unsigned)h.bytes;
                --in_flight;
                if (!pk_walk_fast(ring_s + off, h.n_comp, h.n_click, h.n_conv, h.n_rev, h.bid_cents, h.ctr, h.cvr, lane, lt, r))
                    r = pk_walk_generic(ring + off, h.bytes, h.bid_cents, h.ctr, h.cvr, lane);
            } else {
                r = pk_walk_generic(h.src, h.bytes, h.bid_cents, h.ctr, h.cvr, lane);
            }
            if (lane == 0) {  // hand the sums to the owner lane through the unit's slot
                PkOut o;
                o.I = r.I; o.B = r.B; o.S = r.S; o.overrun = r.overrun; o.cost = r.cost; o.rev = r.rev;
                *reinterpret_cast<PkOut *>(&units[b]) = o;
            }
        }
        __syncwarp();
        int I = 0, B = 0, S = 0;
        long long cost = 0, rev = 0;
        if (pu.bytes > 0) {
            const PkOut o = *reinterpret_cast<const PkOut *>(&units[lane]);
            I = o.I; B = o.B; S = o.S; cost = o.cost; rev = o.rev; my_overrun = o.overrun != 0;
        }

        // ---------------- outputs (coalesced), env completion ----------------
        int safe = 0;
        if (valid) {
            a.out.impressions[u] = I;
            a.out.clicks[u] = B;
            a.out.conversions[u] = S;
            a.out.cost_cents[u] = cost;
            a.out.revenue_cents[u] = rev;
            store_f(a.out.cost, a.out.float_dtype, u, cents_to_dollars(cost));
            store_f(a.out.revenue, a.out.float_dtype, u, cents_to_dollars(rev));
            store_flat_unit(a, e, (int)(u - (int64_t)e * K), I, B, S, cents_to_dollars(cost), cents_to_dollars(rev));
            safe = unit_done(a, e, rev - cost, cost, my_overrun);
        }
        if (a.drift.mask != nullptr || a.out.episode_profit_cents != nullptr) {
            unsigned dm = __ballot_sync(FULL, safe != 0);
            while (dm) {
                const int src = __ffs(dm) - 1;
                dm &= dm - 1;
                const int ee = __shfl_sync(FULL, e, src);
                episode_accumulate(a, ee, lane, 32);
                if (a.drift.mask == nullptr) continue;
                for (int kk = lane; kk < K; kk += 32) {
                    if (!drift_wanted(a, kk)) continue;
                    drift_apply(a, ee, kk, unit_drift<TapeSrc>(a, &t, ee, kk, make_uint4(0, 0, 0, 0)));
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// generic kernel: one thread per unit, lanes in order, no budget
// ------------------------------------------------------------------------------------------
template <typename Src, bool kExplicit>
__global__ void __launch_bounds__(128)
adc_units_kernel(const __grid_constant__ adc_step_args a, const __grid_constant__ adc_tape tape)
{
    const int K = a.kw.K;
    const int64_t total = (int64_t)a.E * K;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) reset_next_counters(a);
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += stride) {
        const int e = (int)(u / K);
        const int k = (int)(u - (int64_t)e * K);
        Src src;
        if constexpr (Src::kTape) {
            src.t = &tape;
        } else {
            src.k0 = (uint32_t)a.seed;
            src.k1 = (uint32_t)(a.seed >> 32);
            src.step = a.step;
            src.env = philox_env(a, e);
        }
        const UnitPar p = load_unit_par(a, e, k);
        uint4 uw;
        const long long V = unit_volume(a, src, &tape, e, k, &uw);
        const long long q = V / ADC_SUBSTEPS;
        UnitCur cur = {0, 0, 0, 0, 0};
        int I = 0, B = 0, S = 0;
        long long cost_c = 0, rev_c = 0;
        double cost_f = 0.0, binf = 0.0;
        bool overrun = false;
        for (int t = 0; t < ADC_SUBSTEPS; ++t) {
            const long long n = t == 0 ? V - (ADC_SUBSTEPS - 1) * q : q;  // bsim:151-167
            const LaneOut o =
                lane_walk<Src, kExplicit, false>(src, &tape, u, k, t, n, p, cur, binf, cost_f);
            I += o.I;
            B += o.B;
            S += o.S;
            cost_c += o.cost_cents;
            rev_c += o.rev_cents;
            overrun |= o.overrun;
        }
        // a truncated tape means the recorded run stopped early: make the env look unaffordable
        // so that unit_done queues it for the exact serial walk (2^40 cents per unit: K < 2^20
        // of them still fit the int64 env accumulator)
        a.out.impressions[u] = I;
        a.out.clicks[u] = B;
        a.out.conversions[u] = S;
        a.out.revenue_cents[u] = rev_c;
        store_f(a.out.revenue, a.out.float_dtype, u, cents_to_dollars(rev_c));
        int safe;
        if constexpr (kExplicit) {
            a.out.cost_cents[u] = 0;
            a.scratch.unit_cost_f64[u] = overrun ? __longlong_as_double(0x7FF0000000000000LL) : cost_f;
            store_f(a.out.cost, a.out.float_dtype, u, cost_f);
            store_flat_unit(a, e, k, I, B, S, cost_f, cents_to_dollars(rev_c));
            safe = unit_done(a, e, 0, 0);
        } else {
            a.out.cost_cents[u] = cost_c;
            store_f(a.out.cost, a.out.float_dtype, u, cents_to_dollars(cost_c));
            store_flat_unit(a, e, k, I, B, S, cents_to_dollars(cost_c), cents_to_dollars(rev_c));
            safe = unit_done(a, e, rev_c - cost_c, cost_c, overrun);
        }
        if (safe) episode_accumulate(a, e, 0, 1);
        if (safe && a.drift.mask != nullptr) {
            for (int kk = 0; kk < K; ++kk) {
                if (!drift_wanted(a, kk)) continue;
                uint4 w = make_uint4(0, 0, 0, 0);
                if constexpr (!Src::kTape) w = src.draw(ST_UNIT, (uint32_t)kk, 0u);
                drift_apply(a, e, kk, unit_drift<Src>(a, &tape, e, kk, w));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// exact serial kernel: one thread per queued env, (sub-step, keyword, click) order, shared budget
// ------------------------------------------------------------------------------------------
// Running day counts of the exact serial kernels: in the outputs themselves, or -- when the
// caller provides adc_scratch.acc_* (outputs in mapped host memory, where a read-modify-write
// would cross PCIe every sub-step) -- in device scratch, stored to the outputs once per env.
struct SerCounts {
    int32_t *I, *B, *S;
};

__device__ __forceinline__ SerCounts ser_counts(const adc_step_args &a)
{
    SerCounts c;
    const bool own = a.scratch.acc_impressions != nullptr;
    c.I = own ? a.scratch.acc_impressions : a.out.impressions;
    c.B = own ? a.scratch.acc_clicks : a.out.clicks;
    c.S = own ? a.scratch.acc_conversions : a.out.conversions;
    return c;
}

__device__ __forceinline__ void ser_publish(const adc_step_args &a, const SerCounts &c, int64_t u)
{
    if (c.I != a.out.impressions) {
        a.out.impressions[u] = c.I[u];
        a.out.clicks[u] = c.B[u];
        a.out.conversions[u] = c.S[u];
    }
}

template <typename Src>
__global__ void __launch_bounds__(64)
adc_serial_kernel(const __grid_constant__ adc_step_args a, const __grid_constant__ adc_tape tape)
{
    const int K = a.kw.K;
    const int count = a.scratch.serial_count[a.parity & 1u];
    const SerCounts acc = ser_counts(a);
    const bool explicit_kw = a.kw.kind != ADC_IMPLICIT;  // un-rounded f64 costs (explicit / multi-bidder keywords)
    const int stride = gridDim.x * blockDim.x;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += stride) {
        const int e = a.scratch.serial_list[idx];
        Src src;
        if constexpr (Src::kTape) {
            src.t = &tape;
        } else {
            src.k0 = (uint32_t)a.seed;
            src.k1 = (uint32_t)(a.seed >> 32);
            src.step = a.step;
            src.env = philox_env(a, e);
        }
        for (int k = 0; k < K; ++k) {
            const int64_t u = (int64_t)e * K + k;
            acc.I[u] = 0;
            acc.B[u] = 0;
            acc.S[u] = 0;
            a.out.cost_cents[u] = 0;
            a.out.revenue_cents[u] = 0;
            if (explicit_kw) a.scratch.unit_cost_f64[u] = 0.0;
            if (a.detail.costs != nullptr) {
                a.detail.n_recorded[u] = 0;
                a.detail.volume_seen[u] = 0.0;
                for (int t = 0; t < ADC_SUBSTEPS; ++t) {
                    a.detail.lane_clicks[u * ADC_SUBSTEPS + t] = 0;
                    a.detail.lane_convs[u * ADC_SUBSTEPS + t] = 0;
                }
            }
        }
        const adc_detail *det = a.detail.costs != nullptr ? &a.detail : nullptr;
        const double budget = step_budget(a, e);
        double remaining = budget;  // bsim:214
        bool stop = false;
        for (int t = 0; t < ADC_SUBSTEPS && !stop; ++t) {
            for (int k = 0; k < K; ++k) {
                const int64_t u = (int64_t)e * K + k;
                UnitPar p = load_unit_par(a, e, k, !Src::kTape);
                uint4 uw;
                const long long V = unit_volume(a, src, &tape, e, k, &uw);
                p.volume = V;
                const long long q = V / ADC_SUBSTEPS;
                const long long n0 = V - (ADC_SUBSTEPS - 1) * q;
                const long long n = t == 0 ? n0 : q;
                UnitCur cur;
                cur.auction = t == 0 ? 0 : n0 + (long long)(t - 1) * q;
                cur.n_clk = 0;
                if constexpr (!Src::kTape) {
                    if (!explicit_kw && n > 0) cur.n_clk = clicks_before(src, k, p.u2, V, cur.auction);
                }
                cur.n_conv = acc.B[u];
                cur.n_rev = acc.S[u];
                cur.n_cost = acc.I[u];
                cur.n_click = acc.I[u];
                if (a.kw.kind == ADC_EXPLICIT) {
                    if constexpr (Src::kTape) {  // one slot per impression, or one phantom slot
                        int s = 0;
                        for (int tt = 0; tt < t; ++tt) {
                            const int i = tape.impr[u * ADC_SUBSTEPS + tt];
                            s += i < 1 ? 1 : i;
                        }
                        cur.n_click = s;
                    }
                }
                double b = remaining;
                LaneOut o;
                if (explicit_kw) {
                    double day_cost = a.scratch.unit_cost_f64[u];
                    o = lane_walk<Src, true, true>(src, &tape, u, k, t, n, p, cur, b, day_cost, det);
                    a.scratch.unit_cost_f64[u] = day_cost;
                } else {
                    double unused = 0.0;
                    o = lane_walk<Src, false, true>(src, &tape, u, k, t, n, p, cur, b, unused, det);
                }
                if (det != nullptr) {
                    const int nb = acc.B[u] + o.B;
                    det->n_recorded[u] = nb < det->cap ? nb : det->cap;
                    if (o.I >= 1) det->volume_seen[u] += (double)n;  // bsim:130-137
                    det->lane_clicks[u * ADC_SUBSTEPS + t] = o.B;
                    det->lane_convs[u * ADC_SUBSTEPS + t] = o.S;
                }
                acc.I[u] += o.I;
                acc.B[u] += o.B;
                acc.S[u] += o.S;
                a.out.cost_cents[u] += o.cost_cents;
                a.out.revenue_cents[u] += o.rev_cents;
                if (a.budget_alias) remaining = b;                    // bsim:102 on an aliased ndarray
                remaining = __dsub_rn(remaining, o.lane_cost_sum);   // bsim:225
                if (remaining <= 0) {                                // bsim:230-233
                    stop = true;
                    break;
                }
            }
        }
        double reward = 0.0;
        long long profit_c = 0;
        for (int k = 0; k < K; ++k) {
            const int64_t u = (int64_t)e * K + k;
            const double rv = cents_to_dollars(a.out.revenue_cents[u]);
            store_f(a.out.revenue, a.out.float_dtype, u, rv);
            ser_publish(a, acc, u);
            if (explicit_kw) {
                const double c = a.scratch.unit_cost_f64[u];
                store_f(a.out.cost, a.out.float_dtype, u, c);
                store_flat_unit(a, e, k, acc.I[u], acc.B[u], acc.S[u], c, rv);
                reward = __dadd_rn(reward, __dsub_rn(rv, c));
            } else {
                const double c = cents_to_dollars(a.out.cost_cents[u]);
                store_f(a.out.cost, a.out.float_dtype, u, c);
                store_flat_unit(a, e, k, acc.I[u], acc.B[u], acc.S[u], c, rv);
                profit_c += a.out.revenue_cents[u] - a.out.cost_cents[u];
            }
            if (a.out.episode_profit_cents != nullptr)
                a.out.episode_profit_cents[u] += a.out.revenue_cents[u] - a.out.cost_cents[u];
        }
        if (!explicit_kw) reward = cents_to_dollars(profit_c);
        env_tail(a, e, reward, budget, remaining);
        if (a.drift.mask != nullptr) {
            for (int kk = 0; kk < K; ++kk) {
                if (!drift_wanted(a, kk)) continue;
                uint4 w = make_uint4(0, 0, 0, 0);
                if constexpr (!Src::kTape) w = src.draw(ST_UNIT, (uint32_t)kk, 0u);
                drift_apply(a, e, kk, unit_drift<Src>(a, &tape, e, kk, w));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// exact serial kernel, warp-cooperative (free-running implicit keywords): one warp per queued env.
//
// The reference walks (sub-step, keyword, click) with ONE shared float budget (bsim:214-233,
// :97-104), but only the affordability tests are sequential: what is clicked, at what price and
// whether it converts does not depend on the budget.  So a warp
//   (0) EXPANDS the env's day in the hot kernel's form into a slab of global memory it owns
//       (adc_scratch.serial_ws): lane <-> keyword, the bit-sliced outcome masks of every 32-auction
//       group; from them one HEADER WORD per (sub-step, keyword) lane -- impressions, clicked slots,
//       where the lane's slots start in the keyword's slot pool -- laid out sub-step-major (a running
//       popcount at the sub-step ends that fall into the group), and the pool itself: 64 words per
//       keyword, one per clicked slot of the day in click order (price in cents | converts << 31),
//       prices drawn four per Philox call, flattened over the 32 keywords;
//   (1) WALKS: per sub-step and chunk of 32 keywords one coalesced load brings the lanes' headers, a
//       short gather their slots; the whole warp runs ONE uniform scan over them in keyword order --
//       the reference's f64 sequence `if budget >= cost: budget -= cost`, alias rule and
//       `remaining <= 0` exit included -- and a lane whose accepted count differs from its slot
//       count writes it back into its header (nothing else is stored in the loop);
//   (2) COMMITS: lane <-> keyword again, every keyword folds its lanes that ran into a 64-bit mask of
//       the paid slots: conversions are one popcount against the conversion-by-rank bits, prices a
//       masked sum over the pool, then one revenue per conversion by conversion rank -- and stores
//       the day's outputs once.
// Nothing is re-drawn per sub-step and no keyword count is special.  Lanes the slab cannot
// describe (volume > 512, more than 61 impressions or clicked slots in one sub-step, a chunk of 32
// keywords with more than 4096 clicked slots) are
// walked again by lane_walk with the budget (serial_direct_lane) and committed on the spot.
// ------------------------------------------------------------------------------------------
constexpr int kSerWarps = 4;
constexpr int kSlabGroups = 16;    // 32-auction groups per unit and day the slab describes (volume <= 512)
constexpr int kUnitSlots = 512;    // clicked slots of one unit's day the slab describes (every day of <= 512 auctions)
constexpr int kPoolPerUnit = 128;  // a chunk of 32 keywords shares a pool of >= 32 x 128 clicked slots: the smallest slab
constexpr int kMaxPoolPerUnit = kUnitSlots;  // ... and of 32 x 512 when the workspace has the room (no chunk can overflow then)
constexpr int kSerRegs = 6;       // clicked slots of a lane and sub-step the walk holds in registers
constexpr int kSerMinBlocks = 7;   // 28 warps per SM: a 4096-env queue is resident in one wave

// lane header: bits 0..5 clicked slots (after the walk: accepted ones), 6..11 impressions, 12..31 first slot
// in the chunk's pool.  Slot field 63: the lane is walked by lane_walk; 62: it was, and bits 12..31 hold its
// conversions.
constexpr uint32_t kHdrDirect = 63u, kHdrDirectDone = 62u, kHdrMaxCount = 61u;

// slab of one warp: uint32 hdr[24][Kp] (Kp = K rounded up to 32) | uint32 pool[Kp / 32][32 x slots per keyword] |
// uint4 acc[Kp]: per keyword the day's impressions | bit 31: mixed commit, and the running sums of the walk
// (clicks, conversions, cents) | uint32 rmin[24 x Kp / 32]: the round index of the walk
__host__ __device__ inline int64_t slab_kp(int K) { return ((int64_t)K + 31) & ~(int64_t)31; }
__host__ __device__ inline int64_t slab_bytes_of(int K)
{
    return (int64_t)ADC_SUBSTEPS * slab_kp(K) * 4 + slab_kp(K) * kPoolPerUnit * 4 + slab_kp(K) * 16 +
           ((ADC_SUBSTEPS * (slab_kp(K) / 32) * 4 + 15) & ~(int64_t)15);
}
// ... of which everything but the pool:
__host__ __device__ inline int64_t slab_fixed_bytes_of(int K)
{
    return slab_bytes_of(K) - slab_kp(K) * kPoolPerUnit * 4;
}

// The bits of x at the set positions of m, packed towards bit 0 (parallel suffix compress, Hacker's
// Delight 7-4): bit r of the result = x at the r-th set bit of m.
__device__ __forceinline__ uint32_t compress32(uint32_t x, uint32_t m)
{
    x &= m;
    uint32_t mk = ~m << 1;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        uint32_t mp = mk ^ (mk << 1);
        mp ^= mp << 2; mp ^= mp << 4; mp ^= mp << 8; mp ^= mp << 16;
        const uint32_t mv = mp & m;
        m = (m ^ mv) | (mv >> (1 << i));
        const uint32_t t = x & mv;
        x = (x ^ t) | (t >> (1 << i));
        mk &= ~mp;
    }
    return x;
}

// A (sub-step, keyword) lane the slab cannot describe: walked again with the budget by lane_walk.  Out
// of line: its registers stay out of the scan loop.
struct DirectOut {
    int I, B, S;
    long long cost_c, rev_c;
    double next;  // the campaign's remaining budget after this lane (bsim:102 alias, :225)
};

__device__ __noinline__ DirectOut serial_direct_lane(const adc_step_args &a, const PhiloxSrc &src, int e, int k, int t,
                                                     int n_rev, double remaining)
{
    const adc_tape *no_tape = nullptr;
    const int64_t u = (int64_t)e * a.kw.K + k;
    UnitPar p = load_unit_par(a, e, k, true);
    uint4 uw;
    p.volume = unit_volume(a, src, no_tape, e, k, &uw);
    const long long q = p.volume / ADC_SUBSTEPS, n0 = p.volume - (ADC_SUBSTEPS - 1) * q;
    const long long n = t == 0 ? n0 : q, j0 = t == 0 ? 0 : n0 + (long long)(t - 1) * q;
    const int n_clk0 = clicks_before(src, k, p.u2, p.volume, j0);  // price-draw rank of the lane's first click
    double b = remaining, unused = 0.0;
    UnitCur cur = {j0, 0, 0, n_rev, 0, n_clk0};
    const LaneOut o = lane_walk<PhiloxSrc, false, true>(src, no_tape, u, k, t, n, p, cur, b, unused);
    DirectOut d;
    d.I = o.I; d.B = o.B; d.S = o.S; d.cost_c = o.cost_cents; d.rev_c = o.rev_cents;
    d.next = __dsub_rn(a.budget_alias ? b : remaining, o.lane_cost_sum);
    return d;
}

// conversions a keyword's lanes accepted in the sub-steps before t (the revenue-draw rank of the next one)
__device__ __noinline__ int conversions_before(const uint32_t *hdr, const uint32_t *pool, int64_t Kp, int k, int t)
{
    int s = 0;
    for (int tt = 0; tt < t; ++tt) {
        const uint32_t h = hdr[(int64_t)tt * Kp + k];
        const uint32_t f = h & 63u;
        if (f == kHdrDirectDone) { s += (int)(h >> 12); continue; }
        if (f == kHdrDirect) continue;
        for (uint32_t i = 0; i < f; ++i) s += (int)(pool[(h >> 12) + i] >> 31);
    }
    return s;
}

// Commit of a keyword that had lanes walked by lane_walk (rare): the slab lanes that ran, in sub-step
// order, with the revenue-draw ranks the direct lanes' conversions shift.  Out of line.
struct CommitOut {
    int I, B, S;
    long long cost_c, rev_c;
};

__device__ __noinline__ CommitOut commit_mixed_unit(const PhiloxSrc &src, const uint32_t *hdr, const uint32_t *pool,
                                                    int64_t Kp, int k, int t_stop, int k_stop, float rm, float rs)
{
    CommitOut o;
    o.I = o.B = o.S = 0; o.cost_c = o.rev_c = 0;
    uint4 rw = make_uint4(0, 0, 0, 0);
    int rw_idx = -1;
    for (int t = 0; t < ADC_SUBSTEPS; ++t) {
        if (t > t_stop || (t == t_stop && k > k_stop)) break;  // after the early break nothing ran
        const uint32_t h = hdr[(int64_t)t * Kp + k];
        const uint32_t f = h & 63u;
        if (f == kHdrDirectDone) { o.S += (int)(h >> 12); continue; }  // committed on the spot
        if (f == kHdrDirect) continue;
        o.I += (int)((h >> 6) & 63u);
        o.B += (int)f;
        for (uint32_t i = 0; i < f; ++i) {
            const uint32_t w = pool[(h >> 12) + i];
            o.cost_c += (int)(w & 0x7FFFFFFFu);
            if (w >> 31) {  // one revenue per conversion, by conversion rank (direct lanes' included)
                const int r = o.S++;
                if ((r >> 2) != rw_idx) {
                    rw_idx = r >> 2;
                    rw = src.draw(ST_REVENUE, (uint32_t)k, (uint32_t)rw_idx);
                }
                const uint32_t ww = (r & 3) == 0 ? rw.x : (r & 3) == 1 ? rw.y : (r & 3) == 2 ? rw.z : rw.w;
                o.rev_c += revenue_cents(ww, rm, rs);
            }
        }
    }
    return o;
}

template <bool kPrefix>
__global__ void __launch_bounds__(kSerWarps * 32, kSerMinBlocks)
adc_serial_warp_implicit_kernel(const __grid_constant__ adc_step_args a, int n_slabs, long long slab_stride, int pool_per_unit)
{
    // phase 0 stages a chunk's 24 x 32 headers and 8 x 32 conversion-by-rank words here
    __shared__ uint32_t s_stage[kSerWarps][(ADC_SUBSTEPS + kUnitSlots / 32) * 32];
    __shared__ FlatCost s_cost[kSerWarps][32];
    __shared__ __align__(16) double s_lsum[kSerWarps][32];
    __shared__ int s_start[kSerWarps][32];
    __shared__ int s_poff[kSerWarps][32];
    __shared__ unsigned char s_nzl[kSerWarps][32];
    __shared__ float2 s_tab[128];  // Exp(1) sampler table, staged from global
    if (threadIdx.x < 128) s_tab[threadIdx.x] = kNeglogTab[threadIdx.x];
    __syncthreads();
    const int K = a.kw.K;
    const int64_t Kp = slab_kp(K);
    const SerCounts acc = ser_counts(a);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * kSerWarps + warp;
    const int n_warps = min(gridDim.x * kSerWarps, n_slabs);
    const int count = a.scratch.serial_count[a.parity & 1u];
    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const adc_tape *no_tape = nullptr;
    if (gwarp >= n_warps) return;
    unsigned char *const slab_raw = reinterpret_cast<unsigned char *>(a.scratch.serial_ws) + (size_t)gwarp * (size_t)slab_stride;
    const uint32_t ppu = (uint32_t)pool_per_unit;  // clicked slots per keyword in a chunk's pool (a multiple of 4)
    uint32_t *const hdr = reinterpret_cast<uint32_t *>(slab_raw);
    uint32_t *const pool = hdr + (size_t)ADC_SUBSTEPS * Kp;  // chunk c0 / 32 owns pool[c0 * kPoolPerUnit ..)
    uint4 *const acc4 = reinterpret_cast<uint4 *>(pool + (size_t)Kp * ppu);
    uint32_t *const rmin = reinterpret_cast<uint32_t *>(acc4 + Kp);  // [24 x Kp / 32] cheapest first click of a round
    const uint32_t Kp32 = (uint32_t)Kp;
    const int n_chunks = (int)(Kp32 >> 5), n_rounds = ADC_SUBSTEPS * n_chunks;
    uint32_t *const s_hdr = &s_stage[warp][0];            // [24][32]
    uint32_t *const s_cv = s_hdr + ADC_SUBSTEPS * 32;     // [kUnitSlots / 32][32]: bit r of a lane's words: its r-th click converts

    for (int idx = gwarp; idx < count; idx += n_warps) {
        const int e = a.scratch.serial_list[idx];
        PhiloxSrc src{k0, k1, a.step, philox_env(a, e)};
        // ---- phase 0: expand the day into the slab, 32 keywords at a time
        bool env_direct = false;  // some lane of the env is walked by lane_walk: the headers must track the walk
        for (int c0 = 0; c0 < K; c0 += 32) {
            // (a) lane <-> keyword: thresholds, volume, outcome masks -> lane headers, conversion-by-rank bits
            const int k = c0 + lane;
            const bool act = k < K;
            bool whole_direct = false, lane_direct = false;
            int V = 0;
            Unit2 u2;
            u2.t1 = u2.t2 = u2.t3 = u2.full = u2.h1 = u2.a1 = u2.a2 = 0u; u2.L = 0.f; u2.b = 0.f; u2.W = 1;
            int floor_c = 0;
            __syncwarp();
#pragma unroll
            for (int w = 0; w < kUnitSlots / 32; ++w) s_cv[w * 32 + lane] = 0u;
            if (act) {
                const int64_t u = (int64_t)e * K + k;
                acc.I[u] = 0;  // the direct lanes accumulate here; everything else is stored once in phase 2
                acc.B[u] = 0;
                acc.S[u] = 0;
                a.out.cost_cents[u] = 0;
                a.out.revenue_cents[u] = 0;
                const UnitPar p = load_unit_par(a, e, k, true);
                uint4 uw;
                const long long Vl = unit_volume(a, src, no_tape, e, k, &uw);
                const bool beats_rivals = p.u2.W > p.floor_cents;
                // what the slab cannot describe: long days, bids beyond the 16-bit prices
                whole_direct = (Vl > 32 * kSlabGroups || p.win_cents > kMaxFlatBidCents) && beats_rivals;
                V = whole_direct || !beats_rivals ? 0 : (int)Vl;
                u2 = p.u2;
                floor_c = max(p.floor_cents, 0);
            }
            int baseI = 0, baseC = 0;
            bool short_day = false;
            {
                const PhiloxPre pa = philox_pre(a.step, stream_word(ST_AUCTION, 0u, (uint32_t)k), src.env, k0, k1);
                const int q = V / ADC_SUBSTEPS, n0 = V - (ADC_SUBSTEPS - 1) * q;  // bsim:151-167
                int t = 0, end_t = n0;     // sub-step t ends before auction end_t
                int lastI = 0, lastC = 0;  // impressions / clicked slots of the sub-steps already emitted
                const int Gmax = __reduce_max_sync(FULL, (V + 31) >> 5);
                // a chunk of short days (sparse keywords, V < 24): every unit has its whole day in sub-step 0,
                // one header row instead of 24
                short_day = Gmax <= 1 && __all_sync(FULL, V < ADC_SUBSTEPS);
                for (int g = 0; g < Gmax; ++g) {
                    const int rem = V - 32 * g;
                    const uint32_t active = rem >= 32 ? 0xFFFFFFFFu : (rem > 0 ? (1u << rem) - 1u : 0u);
                    const Masks3 m = group_masks(active, (uint32_t)g, u2.t1, u2.t2, u2.t3, u2.full, pa.n0, pa.n1,
                                                 pa.x3, k0, k1);
                    // conversion flags by click rank, appended to the lane's bitmap at its running click count:
                    // the group's conversion bits compressed onto its click bits (a fixed ~80 instructions) when
                    // the slowest lane has many conversions, a loop over them when all have few (sparse keywords)
                    if (__reduce_max_sync(FULL, (unsigned)__popc(m.conv)) > 7u) {
                        const uint32_t cb = compress32(m.conv, m.click);
                        const int w0 = baseC >> 5, sh = baseC & 31;
                        if (w0 < kUnitSlots / 32) s_cv[w0 * 32 + lane] |= cb << sh;
                        if (sh != 0 && w0 + 1 < kUnitSlots / 32) s_cv[(w0 + 1) * 32 + lane] |= cb >> (32 - sh);
                    } else {
                        uint32_t cv = m.conv;
                        while (cv) {
                            const int bpos = __ffs(cv) - 1;
                            cv &= cv - 1;
                            const int r = baseC + __popc(m.click & ((1u << bpos) - 1u));
                            if (r < kUnitSlots) s_cv[(r >> 5) * 32 + lane] |= 1u << (r & 31);
                        }
                    }
                    // the sub-steps that end inside this group: one header each
                    const int hi = min(32 * g + 32, V);
                    while (!short_day && rem > 0 && t < ADC_SUBSTEPS && end_t <= hi) {
                        const int low = end_t - 32 * g;  // 1..32 (0 only for the empty sub-steps of a short day)
                        const uint32_t lm = low >= 32 ? 0xFFFFFFFFu : (1u << low) - 1u;
                        const int cumI = baseI + __popc(m.win & lm), cumC = baseC + __popc(m.click & lm);
                        const int ci = cumI - lastI, cc = cumC - lastC;
                        uint32_t h = (uint32_t)cc | ((uint32_t)ci << 6) | ((uint32_t)min(lastC, kUnitSlots) << 12);
                        if (cc > (int)kHdrMaxCount || ci > (int)kHdrMaxCount) { h = kHdrDirect; lane_direct = true; }  // beyond a header's counts
                        s_hdr[t * 32 + lane] = h;
                        lastI = cumI; lastC = cumC;
                        ++t;
                        end_t += q;
                    }
                    baseI += __popc(m.win);
                    baseC += __popc(m.click);
                }
                // a unit with more clicked slots than the slab describes: every lane is walked by lane_walk
                if (baseC > kUnitSlots) whole_direct = true;
                if (short_day) s_hdr[lane] = (uint32_t)baseC | ((uint32_t)baseI << 6);  // < 24 auctions: the counts fit
                else  // (empty sub-steps after the day's last auction still carry the running first-slot field)
                    for (; t < ADC_SUBSTEPS; ++t) s_hdr[t * 32 + lane] = (uint32_t)min(lastC, kUnitSlots) << 12;
            }
            int B = whole_direct ? 0 : baseC;
            if (act) acc4[k] = make_uint4((uint32_t)baseI, 0u, 0u, 0u);  // the day's impressions | clicks, conversions, cents paid
            {   // room in the chunk's slot pool (units padded to 4 slots); a unit that does not fit is re-walked
                const int b4 = (B + 3) & ~3;
                const int off = warp_incl_scan(b4, lane) - b4;
                if ((uint32_t)(off + b4) > 32u * ppu) { whole_direct = true; B = 0; }
                env_direct = env_direct || __any_sync(FULL, whole_direct || lane_direct);
                s_poff[warp][lane] = off;
                // the lanes' headers, sub-step-major: one coalesced store per sub-step
                if (act) {
                    const uint32_t add = (uint32_t)off << 12;
                    const int rows = short_day ? 1 : ADC_SUBSTEPS;
#pragma unroll 4
                    for (int t = 0; t < rows; ++t) {
                        const uint32_t h = s_hdr[t * 32 + lane];
                        hdr[(uint32_t)t * Kp32 + (uint32_t)k] = whole_direct ? kHdrDirect : ((h & 63u) == kHdrDirect ? h : h + add);
                    }
                    for (int t = rows; t < ADC_SUBSTEPS; ++t)
                        hdr[(uint32_t)t * Kp32 + (uint32_t)k] = whole_direct ? kHdrDirect : (uint32_t)(off + B) << 12;
                }
                const PhiloxPre pc = philox_pre(a.step, stream_word(ST_COST, 0u, (uint32_t)k), src.env, k0, k1);
                FlatCost fc;
                fc.t1 = u2.t1; fc.h1 = u2.h1; fc.a1 = u2.a1; fc.a2 = u2.a2; fc.L = u2.L; fc.b = u2.b;
                fc.W = u2.W | ((u2.full & 1u) ? (int)0x80000000u : 0);
                fc.floor_c = floor_c; fc.n0 = pc.n0; fc.n1 = pc.n1; fc.x3 = pc.x3; fc.B = B;
                s_cost[warp][lane] = fc;
            }
            // (b) one price per clicked slot, 4 per Philox call, flattened over the 32 keywords
            const FlatMap fm = flat_map_begin((B + 3) >> 2, lane, s_start[warp], s_nzl[warp]);
            const int TB = fm.total;
            uint32_t *const cpool = pool + (size_t)c0 * ppu;
            for (int base = 0; base < TB; base += 32) {
                const int b = flat_map_unit(fm, base, lane, s_nzl[warp]);
                const int i = base + lane;
                if (i < TB) {
                    const FlatCost f = s_cost[warp][b];
                    const int q = i - s_start[warp][b];
                    const uint4 w = philox_from_pre((uint32_t)q, f.n0, f.n1, f.x3, k0, k1);
                    const bool t1_full = f.W < 0;
                    const int W = f.W & 0x7FFFFFFF;
                    const uint32_t cvw = s_cv[((4 * q) >> 5) * 32 + b] >> ((4 * q) & 31);  // 4 q .. 4 q + 3 share a word
                    uint4 pk;
                    pk.x = (uint32_t)max(cost_cents2(w.x, f.t1, t1_full, f.h1, f.a1, f.a2, f.L, f.b, W, s_tab), f.floor_c) | ((cvw & 1u) << 31);
                    pk.y = (uint32_t)max(cost_cents2(w.y, f.t1, t1_full, f.h1, f.a1, f.a2, f.L, f.b, W, s_tab), f.floor_c) | (((cvw >> 1) & 1u) << 31);
                    pk.z = (uint32_t)max(cost_cents2(w.z, f.t1, t1_full, f.h1, f.a1, f.a2, f.L, f.b, W, s_tab), f.floor_c) | (((cvw >> 2) & 1u) << 31);
                    pk.w = (uint32_t)max(cost_cents2(w.w, f.t1, t1_full, f.h1, f.a1, f.a2, f.L, f.b, W, s_tab), f.floor_c) | (((cvw >> 3) & 1u) << 31);
                    *reinterpret_cast<uint4 *>(cpool + s_poff[warp][b] + 4 * q) = pk;  // offsets are multiples of 4 slots
                }
            }
            __syncwarp();
            // (c) the round index: the cheapest FIRST click of every (sub-step, chunk) round -- a round whose
            // cheapest first click costs more than what is left of the budget changes nothing (every lane
            // breaks at once, bsim:99-104), and the walk does not visit it.  0: a lane of the round is walked
            // by lane_walk (always visited); all ones: no clicked slot in the round.
            {
                const uint32_t off = (uint32_t)s_poff[warp][lane];
                const int rows = short_day ? 1 : ADC_SUBSTEPS;
                if (short_day) {  // rows 1..23: no clicked slots (or nothing but lane_walk lanes)
                    const bool some_direct = __any_sync(FULL, act && whole_direct);
                    if (lane >= 1 && lane < ADC_SUBSTEPS) rmin[lane * n_chunks + (c0 >> 5)] = some_direct ? 0u : 0xFFFFFFFFu;
                }
#pragma unroll 4
                for (int t = 0; t < rows; ++t) {
                    const uint32_t h = s_hdr[t * 32 + lane];
                    uint32_t v = 0xFFFFFFFFu;
                    if (act) {
                        if (whole_direct || (h & 63u) == kHdrDirect) v = 0u;
                        else if ((h & 63u) != 0u) v = cpool[off + (h >> 12)] & 0x7FFFFFFFu;
                    }
                    v = __reduce_min_sync(FULL, v);
                    if (lane == 0) rmin[t * n_chunks + (c0 >> 5)] = v;
                }
            }
            __syncwarp();
        }
        __threadfence_block();
        __syncwarp();
        // ---- phase 1: the walk.  A lane's clicked slots of the round come into registers with one batch
        // of independent loads (kSerRegs of them; the rare longer lane reads on), the next round's headers are
        // already on their way; what a lane accepted goes straight into its keyword's accumulators in the slab.
        const double budget = step_budget(a, e);
        double remaining = budget;  // warp-uniform (bsim:214)
        bool stop = false;
        bool bound = false;  // some round could not be proven fully affordable: next step comes here directly
        int t_stop = ADC_SUBSTEPS, k_stop = K;  // the lane after which nothing ran (bsim:230-233)
        // Rounds are visited through the index: 32 rounds' cheapest first clicks per load, the next round
        // worth a visit is the first whose value `remaining` still covers.  An env with lane_walk lanes
        // visits every round (their headers track the walk); a budget <= 0 gets its one look at round 0.
        int round = 0;
        // ---- phase 1a: the affordable prefix.  While `remaining` exceeds a round's total spend (exact cents)
        // by more than a cent no `budget >= cost` test of the round can fail: every click is accepted and the
        // walk is `remaining -= lane_sum` lane after lane (bsim:225 + rust sum_list).  Such a round needs
        // nothing but its lanes' own sequential f64 sums: header -> slots -> sums -> the 32-step chain, the next
        // round's headers already on their way; no accumulator, no index, no scan.  The first round that
        // cannot be proven affordable hands over to the general walk below; what the prefix accepted is
        // recounted at the commit from the keywords' slots.  (Alias rule, lane_walk lanes: general walk.)
        int n_prefix = 0;  // rounds [0, n_prefix) were accepted whole
        if (kPrefix && !env_direct && !a.budget_alias && remaining > 0) {
            bool go = true;
            // the headers of a round (this lane's keyword of its chunk) and the chunk's pool base
            auto header_of = [&](int rd, uint32_t &cbase) -> uint32_t {
                const int t_cur = __float2int_rd(__fdividef((float)rd + 0.5f, (float)n_chunks));
                const uint32_t c_cur = (uint32_t)(rd - t_cur * n_chunks) << 5;
                cbase = c_cur * ppu;
                return c_cur + (uint32_t)lane < (uint32_t)K ? hdr[(uint32_t)t_cur * Kp32 + c_cur + (uint32_t)lane] : 0u;
            };
            for (int blk = 0; blk < n_rounds && go; blk += 32) {
                const int r = blk + lane;
                unsigned has = __ballot_sync(FULL, r < n_rounds && rmin[r] != 0xFFFFFFFFu);  // rounds with clicked slots
                n_prefix = min(blk + 32, n_rounds);
                uint32_t cbase = 0u, h = 0u;
                if (has) h = header_of(blk + __ffs(has) - 1, cbase);
                while (has) {
                    const int rd = blk + __ffs(has) - 1;
                    has &= has - 1;
                    const int cc = (int)(h & 63u);
                    const uint32_t *const sp = pool + cbase + (h >> 12);
                    uint32_t ncbase = 0u, nh = 0u;  // the next round's headers are on their way during this round
                    if (has) nh = header_of(blk + __ffs(has) - 1, ncbase);
                    double ls = 0.0;
                    unsigned cents = 0u;
                    const int mcc = (int)__reduce_max_sync(FULL, (unsigned)cc);
                    for (int i0 = 0; i0 < mcc; i0 += kSerRegs) {  // kSerRegs independent loads, then the ordered sum
                        uint32_t w[kSerRegs];
#pragma unroll
                        for (int i = 0; i < kSerRegs; ++i) w[i] = i0 + i < cc ? sp[i0 + i] : 0u;
#pragma unroll
                        for (int i = 0; i < kSerRegs; ++i) {
                            if (i0 + i < cc) {
                                const unsigned c = w[i] & 0x7FFFFFFFu;
                                ls = __dadd_rn(ls, cents32_to_dollars((int)c));
                                cents += c;
                            }
                        }
                    }
                    const unsigned total = __reduce_add_sync(FULL, cents);  // <= 32 lanes x 61 x 65535 < 2^31
                    const double spend = cents32_to_dollars((int)total);
                    if (!(remaining > spend + 0.01)) {
                        go = false;
                        n_prefix = rd;
                        break;
                    }
                    s_lsum[warp][lane] = ls;
                    __syncwarp();
#pragma unroll
                    for (int l = 0; l < 32; l += 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(&s_lsum[warp][l]);
                        remaining = __dsub_rn(__dsub_rn(remaining, v.x), v.y);
                    }
                    __syncwarp();
                    h = nh;
                    cbase = ncbase;
                }
            }
            round = n_prefix;
        }
        uint32_t rm = 0u;  // this lane's round of the current block of 32
        int rm_block = -1;
        while (round < n_rounds && !stop) {
            if (!env_direct && remaining > 0) {
                if ((round >> 5) != rm_block) {
                    rm_block = round >> 5;
                    const int r = (rm_block << 5) + lane;
                    rm = r < n_rounds ? rmin[r] : 0xFFFFFFFFu;
                }
                const bool worth = rm != 0xFFFFFFFFu && remaining >= cents32_to_dollars((int)(rm & 0x7FFFFFFFu));
                const unsigned ahead = 0xFFFFFFFFu << (round & 31);
                const unsigned cand = __ballot_sync(FULL, worth) & ahead;
                // rounds passed over although they have clicked slots: the budget binds
                const unsigned dry = __ballot_sync(FULL, rm != 0xFFFFFFFFu && !worth) & ahead;
                if (cand == 0u) {
                    bound = bound || dry != 0u;
                    round = (rm_block + 1) << 5;
                    continue;
                }
                const int nxt = __ffs(cand) - 1;
                bound = bound || (dry & ((1u << nxt) - 1u)) != 0u;
                round = (rm_block << 5) + nxt;
            }
            // round / n_chunks without the integer-division routine: (round + 0.5) / n is never within 0.5 / n of
            // an integer, far more than the float quotient's error for any slab that fits a header index
            const int t_cur = __float2int_rd(__fdividef((float)round + 0.5f, (float)n_chunks));
            const int c_cur = (round - t_cur * n_chunks) << 5;
            ++round;
            const int k = c_cur + lane;
            const bool act = k < K;
            const uint32_t hoff = (uint32_t)t_cur * Kp32 + (uint32_t)k;
            const uint32_t h = act ? hdr[hoff] : 0u;
            // a chunk without an impression or a clicked slot in this sub-step leaves everything alone
            // (a sparse keyword, V < 24, has its whole day in sub-step 0); `remaining <= 0` on entry still
            // gets its one look
            if (remaining > 0 && __all_sync(FULL, (h & 0xFFFu) == 0u)) continue;
            int nclk = (int)(h & 63u);
            const bool direct = nclk == (int)kHdrDirect;
            const uint32_t my_off = (uint32_t)c_cur * ppu + (h >> 12);
            const uint32_t *const sp = pool + my_off;
            uint32_t w[kSerRegs];
#pragma unroll
            for (int i = 0; i < kSerRegs; ++i) w[i] = (!direct && i < nclk) ? sp[i] : 0u;
            uint4 a4 = make_uint4(0, 0, 0, 0);
            if (act) a4 = acc4[k];
            // nothing affordable -- every lane's first click costs more than `remaining`, so each lane breaks
            // at once (bsim:99-104) and `remaining` does not move: the usual state of the sub-steps after the
            // budget ran dry.  Only the impressions count.
            if (remaining > 0) {
                const bool lane_has = act && nclk > 0;
                const bool none = !lane_has || (!direct && !(remaining >= cents32_to_dollars((int)(w[0] & 0x7FFFFFFFu))));
                if (__all_sync(FULL, none)) {
                    bound = bound || __any_sync(FULL, lane_has);
                    if (act && env_direct && nclk > 0) hdr[hoff] = h & ~63u;
                    continue;
                }
            }
            // the lane's own sequential f64 sum (rust sum_list), exact cents and conversions of all its slots
            unsigned cents = 0, convs = 0;
            double lane_sum = 0.0;
#pragma unroll
            for (int i = 0; i < kSerRegs; ++i) {
                if (!direct && i < nclk) {
                    const unsigned c = w[i] & 0x7FFFFFFFu;
                    lane_sum = __dadd_rn(lane_sum, cents32_to_dollars((int)c));
                    cents += c;
                    convs += w[i] >> 31;
                }
            }
            if (!direct && nclk > kSerRegs) {
                for (int i = kSerRegs; i < nclk; ++i) {
                    const uint32_t ww = sp[i];
                    const unsigned c = ww & 0x7FFFFFFFu;
                    lane_sum = __dadd_rn(lane_sum, cents32_to_dollars((int)c));
                    cents += c;
                    convs += ww >> 31;
                }
            }
            if (direct) nclk = (int)kHdrDirect;  // takes the re-walk turn below
            // ---- one uniform scan over the lanes that have clicks: the reference's budget walk
            int B = direct ? 0 : nclk;
            unsigned todo = __ballot_sync(FULL, act && nclk > 0);
            const bool any_click = todo != 0u;
            if (!(remaining > 0)) todo |= 1u;  // a lane must run for `remaining <= 0` to be seen
            int cutoff = 32;
            DirectOut dout;
            dout.I = dout.B = dout.S = 0; dout.cost_c = dout.rev_c = 0; dout.next = 0.0;
            bool direct_done = false;
            // Everything affordable -- `remaining` exceeds this round's total spend (exact cents, doubled
            // under the alias rule) by more than a cent, so no `budget >= cost` test can fail; without the
            // alias rule the walk is then per lane `remaining -= lane_sum` with lane_sum the lane's own
            // sequential f64 sum (bsim:225 + rust sum_list), which the lanes form in parallel (a lane
            // without clicks subtracts an exact 0.0).
            bool all_accepted = false;
            if (remaining > 0) {
                if (!__any_sync(FULL, direct)) {
                    const unsigned total = __reduce_add_sync(FULL, cents);  // <= 32 lanes x 61 x 65535 < 2^32 (bids capped above)
                    const double spend = cents_to_dollars((long long)total);
                    if (remaining > (a.budget_alias ? spend + spend : spend) + 0.01) {
                        all_accepted = true;
                        if (!a.budget_alias) {
                            s_lsum[warp][lane] = lane_sum;
                            __syncwarp();
#pragma unroll
                            for (int l = 0; l < 32; l += 2) {
                                const double2 v = *reinterpret_cast<const double2 *>(&s_lsum[warp][l]);
                                remaining = __dsub_rn(__dsub_rn(remaining, v.x), v.y);
                            }
                            todo = 0;
                        }
                    }
                }
            }
            bound = bound || (any_click && !all_accepted);
            // From here every lane follows lane l's walk: its slots are read again from the pool, all lanes
            // the same word (the owner's loads above left the lines in L1).
            if (todo) {
                if (all_accepted) {  // alias rule: the lane's own walk already drew on the shared budget (bsim:102)
                    while (todo) {
                        const int l = __ffs(todo) - 1;
                        todo &= todo - 1;
                        double b = remaining;
                        const int n_l = __shfl_sync(FULL, nclk, l);
                        const uint32_t *const sl = pool + __shfl_sync(FULL, my_off, l);
                        for (int i = 0; i < n_l; ++i) b = __dsub_rn(b, cents32_to_dollars((int)(sl[i] & 0x7FFFFFFFu)));
                        remaining = __dsub_rn(b, __shfl_sync(FULL, lane_sum, l));
                    }
                } else {
                    B = 0;  // the scan decides lane by lane; lanes it never reaches accept nothing
                    // A lane whose first click costs more than `remaining` does now cannot afford it later in
                    // the round either (`remaining` only falls): it breaks at once, subtracts an exact 0.0 and
                    // leaves `remaining` as it is -- out of the scan.  After the budget ran dry this leaves the
                    // few lanes with a cheap first click.
                    if (remaining > 0) {
                        const bool cannot = !direct && nclk > 0 && !(remaining >= cents32_to_dollars((int)(w[0] & 0x7FFFFFFFu)));
                        todo &= ~__ballot_sync(FULL, cannot);
                    }
                }
            }
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                const int n_l = __shfl_sync(FULL, nclk, l);
                const uint32_t *const sl = pool + __shfl_sync(FULL, my_off, l);
                double next;
                if (n_l != (int)kHdrDirect) {  // every lane runs lane l's walk on its f64 costs
                    double b = remaining, lsum = 0.0;
                    int nacc = 0;
                    for (int i = 0; i < n_l; ++i) {
                        const double cost = cents32_to_dollars((int)(sl[i] & 0x7FFFFFFFu));
                        if (!(b >= cost)) break;  // bsim:99-104
                        ++nacc;
                        lsum = __dadd_rn(lsum, cost);
                        b = __dsub_rn(b, cost);
                    }
                    if (lane == l) B = nacc;
                    next = __dsub_rn(a.budget_alias ? b : remaining, lsum);  // bsim:102 alias, :225
                } else {  // beyond the slab or the buffer: lane l walks its sub-step again, with the budget
                    next = remaining;
                    if (lane == l) {
                        dout = serial_direct_lane(a, src, e, k, t_cur,
                                                  conversions_before(hdr, pool + (uint32_t)c_cur * ppu, Kp, k, t_cur),
                                                  remaining);
                        direct_done = true;
                        next = dout.next;
                    }
                    next = __shfl_sync(FULL, next, l);
                }
                remaining = next;
                if (remaining <= 0) {  // bsim:230-233
                    stop = true;
                    cutoff = l;
                    t_stop = t_cur;
                    k_stop = c_cur + l;
                    break;
                }
            }
            // the lanes that ran: their keyword's running sums; the header keeps an accepted count below the
            // slot count (the mixed commit and conversions_before read it), a direct lane's totals go to the outputs
            if (act && lane <= cutoff) {
                if (direct) {
                    const int64_t u = (int64_t)e * K + k;
                    if (direct_done) {
                        acc.I[u] += dout.I;
                        acc.B[u] += dout.B;
                        acc.S[u] += dout.S;
                        a.out.cost_cents[u] += dout.cost_c;
                        a.out.revenue_cents[u] += dout.rev_c;
                    }
                    hdr[hoff] = kHdrDirectDone | ((uint32_t)dout.S << 12);  // (never reached: nothing of it counts)
                    a4.x |= 0x80000000u;  // the keyword takes the mixed commit
                } else {
                    if (B != nclk) {
                        if (env_direct) hdr[hoff] = (h & ~63u) | (uint32_t)B;
                        cents = 0; convs = 0;
                        for (int i = 0; i < B; ++i) {
                            const uint32_t ww = sp[i];
                            cents += ww & 0x7FFFFFFFu;
                            convs += ww >> 31;
                        }
                    }
                    a4.y += (uint32_t)B;
                    a4.z += convs;
                    a4.w += cents;
                }
                acc4[k] = a4;
            }
            __syncwarp();
        }
        __threadfence_block();
        __syncwarp();
        // ---- phase 2: commit, lane <-> keyword: one revenue per conversion by conversion rank, the outputs
        long long profit_c = 0;
        for (int k = lane; k < K; k += 32) {
            const int64_t u = (int64_t)e * K + k;
            const int64_t pi = (int64_t)e * a.kw.env_stride + k;
            const float rev_mean = (float)a.kw.rev_mean[pi], rev_sd = (float)a.kw.rev_std[pi];
            const uint4 a4 = acc4[k];
            int I = (int)(a4.x & 0x7FFFFFFFu), B = (int)a4.y, S = (int)a4.z;
            long long cost_c = a4.w, rev_c = 0;
            if (n_prefix > 0) {
                // the keyword's sub-steps inside the affordable prefix: every clicked slot was accepted -- its
                // first t_k sub-steps' slots are one contiguous run of the chunk's pool
                const int t_full = __float2int_rd(__fdividef((float)n_prefix + 0.5f, (float)n_chunks));
                const int t_k = t_full + ((k >> 5) < n_prefix - t_full * n_chunks ? 1 : 0);
                if (t_k > 0) {
                    const uint32_t first = hdr[(uint32_t)k] >> 12;  // sub-step 0: the keyword's first slot
                    // (every header's slot field is the keyword's running slot count, empty sub-steps included)
                    const uint32_t hl = hdr[(uint32_t)(t_k - 1) * Kp32 + (uint32_t)k];
                    const uint32_t n = (hl >> 12) + (hl & 63u) - first;
                    const uint32_t *const sp = pool + (uint32_t)(k & ~31) * ppu + first;
                    unsigned cents = 0u, convs = 0u;
                    for (uint32_t i = 0; i < n; ++i) {
                        const uint32_t w = sp[i];
                        cents += w & 0x7FFFFFFFu;
                        convs += w >> 31;
                    }
                    B += (int)n;
                    S += (int)convs;
                    cost_c += cents;
                }
            }
            if (!(a4.x >> 31)) {
                if (stop) {  // the lanes after the early break never ran: only the others' impressions count
                    I = 0;
                    for (int t = 0; t < ADC_SUBSTEPS; ++t) {
                        if (t > t_stop || (t == t_stop && k > k_stop)) break;
                        I += (int)((hdr[(uint32_t)t * Kp32 + (uint32_t)k] >> 6) & 63u);
                    }
                }
                for (int r4 = 0; 4 * r4 < S; ++r4) {
                    const uint4 rw = src.draw(ST_REVENUE, (uint32_t)k, (uint32_t)r4);
                    const int left = S - 4 * r4;
                    rev_c += revenue_cents(rw.x, rev_mean, rev_sd);
                    if (left > 1) rev_c += revenue_cents(rw.y, rev_mean, rev_sd);
                    if (left > 2) rev_c += revenue_cents(rw.z, rev_mean, rev_sd);
                    if (left > 3) rev_c += revenue_cents(rw.w, rev_mean, rev_sd);
                }
            } else {
                // a keyword with lanes walked by lane_walk: their conversions shift the revenue ranks of the
                // slab lanes (S of the mixed commit counts them too: they are in acc.S already)
                const CommitOut o = commit_mixed_unit(src, hdr, pool + (uint32_t)(k & ~31) * ppu, Kp, k, t_stop, k_stop,
                                                      rev_mean, rev_sd);
                I = o.I + acc.I[u];
                B = o.B + acc.B[u];
                S = o.S;
                cost_c = o.cost_c + a.out.cost_cents[u];
                rev_c = o.rev_c + a.out.revenue_cents[u];
            }
            acc.I[u] = I;
            acc.B[u] = B;
            acc.S[u] = S;
            a.out.cost_cents[u] = cost_c;
            a.out.revenue_cents[u] = rev_c;
            store_f(a.out.cost, a.out.float_dtype, u, cents_to_dollars(cost_c));
            store_f(a.out.revenue, a.out.float_dtype, u, cents_to_dollars(rev_c));
            if (acc.I != a.out.impressions) {
                a.out.impressions[u] = I;
                a.out.clicks[u] = B;
                a.out.conversions[u] = S;
            }
            store_flat_unit(a, e, k, I, B, S, cents_to_dollars(cost_c), cents_to_dollars(rev_c));
            if (a.out.episode_profit_cents != nullptr) a.out.episode_profit_cents[u] += rev_c - cost_c;
            profit_c += rev_c - cost_c;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) profit_c += __shfl_xor_sync(FULL, profit_c, off);
        if (lane == 0) {
            env_tail(a, e, cents_to_dollars(profit_c), budget, remaining);
            if (a.scratch.serial_hint != nullptr) a.scratch.serial_hint[e] = bound ? 1 : 0;
        }
        if (a.out.rows != nullptr) {
            __threadfence();  // the row reads what this warp's lanes just stored
            __syncwarp();
            pack_row_warp(a, e, reinterpret_cast<unsigned char *>(a.out.rows), lane);
        }
        if (a.out.unit_records != nullptr) {
            __threadfence();
            __syncwarp();
            pack_units_warp(a, e, lane);
        }
        if (a.drift.mask != nullptr) {
            for (int kk = lane; kk < K; kk += 32) {
                if (!drift_wanted(a, kk)) continue;
                const uint4 w = src.draw(ST_UNIT, (uint32_t)kk, 0u);
                drift_apply(a, e, kk, drift_from_words(a, w));
            }
        }
        __syncwarp();
    }
}




// ------------------------------------------------------------------------------------------
// compact host rows (adc_step_host): one warp per env packs the step's observation into one row
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adc_pack_rows_kernel(const __grid_constant__ adc_step_args a, unsigned char *rows)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp; e < a.E; e += n_warps) pack_row_warp(a, (int)e, rows, lane);
}

__global__ void __launch_bounds__(256)
adc_pack_units_kernel(const __grid_constant__ adc_step_args a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp; e < a.E; e += n_warps) pack_units_warp(a, (int)e, lane);
}

// ------------------------------------------------------------------------------------------
// shared auctions (env_group = A > 1): the rows that cannot win.  One thread per (world, keyword):
// the unique top bidder among the world's A bids is left to the hot kernel; every other row's unit
// (and every unit of an env routed by serial_hint) is finished here -- zero outputs, env completion,
// episode accumulation and drift of the envs that complete -- and marked in adc_scratch.outbid_mask.
// Compact on purpose: seven of eight rows of an 8-bidder world end here, and as straight-line
// passages of the hot kernel they made its instruction caches thrash.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adc_outbid_rows_kernel(const __grid_constant__ adc_step_args a)
{
    const int K = a.kw.K, A = a.env_group;
    const int64_t total = (int64_t)(a.E / A) * K;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const int64_t n_thr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane; base < total; base += n_thr) {
        const int64_t idx = base + lane;
        const bool in = idx < total;
        int w = 0, k = 0;
        int top = (int)0x80000000, n_top = 0;
        if (in) {
            w = (int)(idx / K);
            k = (int)(idx - (int64_t)w * K);
#pragma unroll 4
            for (int r = 0; r < A; ++r) {  // (independent loads: four rows' bids in flight)
                const int c = bid_to_cents(load_f(a.bids, a.bids_dtype, ((int64_t)w * A + r) * K + k));
                if (c > top) { top = c; n_top = 1; }
                else if (c == top) ++n_top;
            }
        }
        // the zero outputs of all the rows that end here first, ONE fence, then their completions
        unsigned mine_m = 0u, hint_m = 0u;  // rows of this (world, keyword) finished here / routed by serial_hint
        if (in) {
            for (int r = 0; r < A; ++r) {
                const int e = w * A + r;
                const int64_t u = (int64_t)e * K + k;
                const int c = bid_to_cents(load_f(a.bids, a.bids_dtype, u));
                const bool hinted = a.scratch.serial_hint != nullptr && a.scratch.serial_hint[e] != 0;
                const bool mine = hinted || !(c == top && n_top == 1);  // unit_floor: outbid <=> not the unique top bid
                a.scratch.outbid_mask[u] = mine ? 1 : 0;
                if (!mine) continue;
                mine_m |= 1u << r;
                if (hinted) hint_m |= 1u << r;
                a.out.impressions[u] = 0;
                a.out.clicks[u] = 0;
                a.out.conversions[u] = 0;
                a.out.cost_cents[u] = 0;
                a.out.revenue_cents[u] = 0;
                store_f(a.out.cost, a.out.float_dtype, u, 0.0);
                store_f(a.out.revenue, a.out.float_dtype, u, 0.0);
                store_flat_unit(a, e, k, 0, 0, 0, 0.0, 0.0);
                if (a.out.rows != nullptr) pack_row_unit(a, e, k, 0, 0, 0, 0.0, 0.0);
                if (a.out.unit_records != nullptr) store_unit_record(a, u, 0, 0, 0, 0.0, 0.0);
            }
        }
        __threadfence();
        for (int r = 0; r < A; ++r) {  // (warp-uniform: the env tails below are warp-cooperative)
            int safe = 0;
            const int e = w * A + r;
            // the warp's units of one env (consecutive keywords of one world's row) are published by one lane
            const bool mine = in && ((mine_m >> r) & 1u);
            const unsigned peers = __match_any_sync(FULL, mine ? w : -1 - lane);
            if (mine && lane == __ffs(peers) - 1) {
                safe = unit_done(a, e, 0, 0, ((hint_m >> r) & 1u) != 0, true, __popc(peers));
                if (safe && a.out.rows != nullptr) pack_row_tail(a, e);
            }
            if (a.drift.mask != nullptr || a.out.episode_profit_cents != nullptr) {
                unsigned todo = __ballot_sync(FULL, safe != 0);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int ee = __shfl_sync(FULL, e, src);
                    episode_accumulate(a, ee, lane, 32);
                    if (a.drift.mask == nullptr) continue;
                    const uint32_t ge = philox_env(a, ee);
                    for (int kk = lane; kk < K; kk += 32) {
                        if (!drift_wanted(a, kk)) continue;
                        const uint4 wd = philox4x32_10(0u, a.step, stream_word(ST_UNIT, 0u, (uint32_t)kk), ge, k0, k1);
                        drift_apply(a, ee, kk, drift_from_words(a, wd));
                    }
                }
            }
            __syncwarp();
        }
    }
}

__global__ void adc_reset_envs_kernel(int32_t E, const uint8_t *mask, double *cum_profit, int32_t *day)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E && (mask == nullptr || mask[e])) {
        cum_profit[e] = 0.0;
        day[e] = 0;
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_pack_rows(const adc_step_args &a, void *rows_dev, cudaStream_t s, int64_t *launches);

static int g_num_sms = 0;

static int num_sms()
{
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int W, int RING, int DEPTH, int MINB>
static cudaError_t launch_packed(const adc_step_args &a, const adc_tape &tp, cudaStream_t s, int64_t *launches)
{
    auto kern = adc_replay_packed_kernel<W, RING, DEPTH, MINB>;
    constexpr int block = W * 32;
    constexpr size_t dyn = (size_t)W * RING + 640;  // (pk_walk_fast's unconditional loads may pass a record's end by 512 + 32 bytes)
    static bool configured = false;
    if (!configured) {
        const cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    const int64_t total = (int64_t)a.E * a.kw.K;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, dyn);
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms() * per_sm;
    const int64_t want = ((total + 31) / 32 + W - 1) / W;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, block, dyn, s>>>(a, tp);
    ++*launches;
    return cudaGetLastError();
}

template <typename K>
static int64_t grid_for(K kernel, int block, int64_t work_items)
{
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms() * per_sm;
    const int64_t want = (work_items + block - 1) / block;
    if (want < grid) grid = want;
    return grid < 1 ? 1 : grid;
}

cudaError_t launch_step(const adc_step_args &a, const adc_tape *tape, cudaStream_t s, int64_t *launches)
{
    const bool explicit_kw = a.kw.kind != ADC_IMPLICIT;  // thread-per-unit kernels, un-rounded f64 costs
    const int64_t total = (int64_t)a.E * a.kw.K;
    cudaError_t err = cudaSuccess;
    adc_tape t0 = {};
    const adc_tape &tp = tape ? *tape : t0;
    // the warp-cooperative exact walk needs its workspace; it is also the only consumer of serial_hint
    const int64_t slab_bytes = slab_bytes_of(a.kw.K);
    const int64_t n_slabs = a.scratch.serial_ws != nullptr ? a.scratch.serial_ws_bytes / slab_bytes : 0;
    const bool warp_walk = tape == nullptr && !explicit_kw && a.n_lanes != 1 && a.detail.costs == nullptr && n_slabs > 0;
    if (tape == nullptr && !explicit_kw) {
        const int block = kFlatWarps * 32;
        int per_sm = 0;
        constexpr int kBU = 32;  // units per big batch
        const bool fl = a.floor_cents != nullptr || a.env_group > 1;
        adc_step_args b = a;
        if (!warp_walk) b.scratch.serial_hint = nullptr;  // (only the warp-cooperative walk keeps the marks up to date)
        if (a.env_group > 1 && a.env_group <= 32 && a.floor_cents == nullptr && a.scratch.outbid_mask != nullptr) {
            // shared auctions: the compact pre-pass finishes the rows that cannot win
            const int64_t pairs = (int64_t)(a.E / a.env_group) * a.kw.K;
            const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((pairs + 255) / 256, (int64_t)num_sms() * 16));
            adc_outbid_rows_kernel<<<(unsigned)grid, 256, 0, s>>>(b);
            ++*launches;
            err = cudaGetLastError();
            if (err != cudaSuccess) return err;
        } else {
            b.scratch.outbid_mask = nullptr;  // (an explicit floor table: every unit goes through the hot kernel)
        }
        // (the variant that may spread a batch's (unit, group) pairs over the lanes is 2.4 % slower on sparse
        // keyword sets, which never use it: the caller says which one it wants, adc_step_args.spread_outcomes)
        const bool sp = a.spread_outcomes != 0;
        void (*kern)(adc_step_args) = fl ? (sp ? adc_flat2_implicit_kernel<true, true> : adc_flat2_implicit_kernel<true, false>)
                                         : (sp ? adc_flat2_implicit_kernel<false, true> : adc_flat2_implicit_kernel<false, false>);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, 0);
        if (per_sm < 1) per_sm = 1;
        int64_t grid = (int64_t)num_sms() * per_sm;
        const int64_t want = ((total + kBU - 1) / kBU + kFlatWarps - 1) / kFlatWarps;
        if (want < grid) grid = want;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, block, 0, s>>>(b);
        ++*launches;
        err = cudaGetLastError();
    } else if (tape == nullptr && a.kw.kind == ADC_EXPLICIT && a.n_lanes != 1) {
        // units per warp and batch: 32 on large steps, fewer on small ones so that every SM gets ~64 warps
        int cnt = 32;
        while (cnt > 1 && total / cnt < (int64_t)num_sms() * 64) cnt >>= 1;
        auto kern = adc_flat_explicit_kernel;
        const int64_t n_batches = (total + cnt - 1) / cnt;
        kern<<<(unsigned)grid_for(kern, kFlatWarps * 32, n_batches * 32), kFlatWarps * 32, 0, s>>>(a, cnt);
        ++*launches;
        err = cudaGetLastError();
    } else if (tape == nullptr) {  // multi-bidder keywords; explicit keywords with n_lanes = 1 (A/B against the flattened kernel)
        auto kern = adc_units_kernel<PhiloxSrc, true>;
        kern<<<(unsigned)grid_for(kern, 128, total), 128, 0, s>>>(a, tp);
        ++*launches;
        err = cudaGetLastError();
    } else if (explicit_kw) {
        auto kern = adc_units_kernel<TapeSrc, true>;
        kern<<<(unsigned)grid_for(kern, 128, total), 128, 0, s>>>(a, tp);
        ++*launches;
        err = cudaGetLastError();
    } else if (a.n_lanes == 1) {  // one thread per unit (kept for A/B against the warp kernel)
        auto kern = adc_units_kernel<TapeSrc, false>;
        kern<<<(unsigned)grid_for(kern, 128, total), 128, 0, s>>>(a, tp);
        ++*launches;
        err = cudaGetLastError();
    } else if (tp.packed != nullptr) {
        // ring geometry: 8 warps x 6 KB per CTA, three CTAs (24 warps) per SM; measured on C2: 6 KB
        // 0.191 ms (5.75 / 6.25 / 6.5 KB: 0.194 / 0.192 / 0.190), 7 KB 0.203 ms, 5 KB with 64 registers
        // and four CTAs 0.205 ms, 8 KB (two CTAs per SM) 0.26 ms.
#ifndef ADC_PK_W
#define ADC_PK_W 8
#define ADC_PK_RING 6144
#define ADC_PK_MINB 3
#endif
        err = launch_packed<ADC_PK_W, ADC_PK_RING, 4, ADC_PK_MINB>(a, tp, s, launches);
    } else {
        auto kern = adc_replay_implicit_kernel;
        kern<<<(unsigned)grid_for(kern, kReplayWarps * 32, total), kReplayWarps * 32, 0, s>>>(a, tp);
        ++*launches;
        err = cudaGetLastError();
    }
    if (err != cudaSuccess) return err;
    // exact serial walk of the queued envs (reads the count on the device; exits at once if 0)
    if (warp_walk) {
        // one warp per queued env, each with its own slab of the workspace
        // The variant with the affordable prefix pays where a day has many rounds the budget covers whole (C2 with
        // budget 1000: 41 of the 44 visited rounds, -20 %); compiled in where it cannot run (alias rule) or
        // rarely does (1000 keywords: 12 of 768 rounds) it costs 2-3 %.  Same results either way.
        const bool prefix = !a.budget_alias && a.kw.K <= 256;
        void (*kern)(adc_step_args, int, long long, int) =
            prefix ? adc_serial_warp_implicit_kernel<true> : adc_serial_warp_implicit_kernel<false>;
        int64_t grid = grid_for(kern, kSerWarps * 32, (int64_t)a.E * 32);
        const int64_t by_ws = (n_slabs + kSerWarps - 1) / kSerWarps;
        if (by_ws < grid) grid = by_ws;
        // The warps that run share the whole workspace: what a slab has beyond the smallest layout goes to its
        // slot pools (up to 512 slots per keyword, where no chunk of keywords can overflow its pool any more).
        const int64_t n_run = std::min<int64_t>(std::min<int64_t>(n_slabs, grid * kSerWarps), 0x7FFFFFFF);
        const int64_t stride = (a.scratch.serial_ws_bytes / n_run) & ~(int64_t)15;
        int64_t ppu = (stride - slab_fixed_bytes_of(a.kw.K)) / (slab_kp(a.kw.K) * 4) & ~(int64_t)3;
        if (ppu > kMaxPoolPerUnit) ppu = kMaxPoolPerUnit;
        kern<<<(unsigned)grid, kSerWarps * 32, 0, s>>>(a, (int)n_run, (long long)stride, (int)ppu);
    } else if (tape == nullptr) {
        auto kern = adc_serial_kernel<PhiloxSrc>;
        kern<<<(unsigned)grid_for(kern, 64, a.E), 64, 0, s>>>(a, tp);
    } else {
        auto kern = adc_serial_kernel<TapeSrc>;
        kern<<<(unsigned)grid_for(kern, 64, a.E), 64, 0, s>>>(a, tp);
    }
    ++*launches;
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    // compact rows: the free-running implicit kernels pack them as they finalise each env; every
    // other kernel family gets one packing pass over the finished step
    if (a.out.unit_records != nullptr && !warp_walk) {
        const int64_t want = ((int64_t)a.E * 32 + 255) / 256;
        const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms() * 8));
        adc_pack_units_kernel<<<(unsigned)grid, 256, 0, s>>>(a);
        ++*launches;
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    if (a.out.rows != nullptr && !warp_walk) return launch_pack_rows(a, a.out.rows, s, launches);
    return cudaSuccess;
}

int64_t serial_slab_bytes(int32_t K) { return slab_bytes_of(K); }

int64_t host_row_bytes(int32_t K, int32_t float_dtype) { return row_layout(K, float_dtype).bytes; }

cudaError_t launch_pack_rows(const adc_step_args &a, void *rows_dev, cudaStream_t s, int64_t *launches)
{
    const int64_t want = ((int64_t)a.E * 32 + 255) / 256;
    int64_t grid = (int64_t)num_sms() * 8;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    adc_pack_rows_kernel<<<(unsigned)grid, 256, 0, s>>>(a, reinterpret_cast<unsigned char *>(rows_dev));
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_reset_envs(int32_t E, const uint8_t *mask, double *cum_profit, int32_t *day,
                              cudaStream_t s, int64_t *launches)
{
    adc_reset_envs_kernel<<<(E + 255) / 256, 256, 0, s>>>(E, mask, cum_profit, day);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace adc
