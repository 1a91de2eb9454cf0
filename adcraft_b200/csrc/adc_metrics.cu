// Ideal-profit estimator of the AKNCP / NCP metrics on the device (sm_100a).
//
// Reference (file:line):
//   get_implicit_kw_bid_cpc_impressions   adcraft/experiment_utils/experiment_metrics.py:20-37
//   get_max_expected_bid_profits          adcraft/experiment_utils/experiment_metrics.py:40-61
// Per keyword the reference samples n = 2048 competitor bids (ImplicitKeyword.sample_bids ->
// bid_abs_laplace, synthetic_kw_helpers.py:104-113), sorts them, and for every bid b of a grid takes
//   idx  = searchsorted(sorted, b, side="right")          # samples <= b
//   rate = idx / n;   cpc = mean(sorted[: min(idx, n-1) + 1])   (the mean includes ONE sample above b)
//   profit(b) = max(vol_mean * rate * bctr * (sctr * mean_rev - cpc), 0)
// and returns max_b profit, the share of grid bids with positive profit, and the argmax.
//
// The samples are whole cents, so the sort is a counting sort: one warp per (env, keyword) unit
// builds a shared-memory histogram over the cents the grid can distinguish, an inclusive scan gives
// count and sum below every cent, and "the next sample above b" is the next non-empty bucket.  All
// sums are exact integers (the reference adds float64 dollars; the two differ by rounding only).
#include "adc_rng.cuh"
#include "adc_step.h"

#include <cmath>

namespace adc {

constexpr int kIdealWarps = 8;
constexpr int kIdealBuckets = 512;  // cents 0 .. 510 exact, bucket 511 = everything above
constexpr int kIdealMaxGrid = ADC_IDEAL_MAX_GRID;

struct IdealGrid {  // by value in the launch: largest cent value c with c / 100 <= bid_grid[g] (-1: none)
    short thr[kIdealMaxGrid];
};

__global__ void __launch_bounds__(kIdealWarps * 32)
adc_ideal_profit_kernel(const __grid_constant__ adc_ideal_args a, const __grid_constant__ IdealGrid grid)
{
    __shared__ unsigned s_cnt[kIdealWarps][kIdealBuckets];  // histogram, then samples with cents <= c
    __shared__ unsigned s_sum[kIdealWarps][kIdealBuckets];  // their sum in cents (<= 2^16 samples x 510)
    __shared__ unsigned s_over_min[kIdealWarps];
    const int K = a.kw.K;
    const int64_t total = (int64_t)a.E * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    unsigned *cnt = s_cnt[warp], *sum = s_sum[warp];
    const int n = a.n_samples;
    for (int64_t u = (int64_t)blockIdx.x * kIdealWarps + warp; u < total; u += (int64_t)gridDim.x * kIdealWarps) {
        const int e = (int)(u / K), k = (int)(u - (int64_t)e * K);
        const int64_t pi = (int64_t)e * a.kw.env_stride + k;
        for (int i = lane; i < kIdealBuckets; i += 32) cnt[i] = 0u;
        if (lane == 0) s_over_min[warp] = 0xFFFFFFFFu;
        __syncwarp();
        // ---- the n sampled competitor bids -> histogram over cents
        auto put = [&](int c) {
            if (c < kIdealBuckets - 1) {
                atomicAdd(&cnt[c], 1u);
            } else {
                atomicAdd(&cnt[kIdealBuckets - 1], 1u);
                atomicMin(&s_over_min[warp], (unsigned)c);
            }
        };
        if (a.samples_cents != nullptr) {
            const int32_t *src = a.samples_cents + u * n;
            for (int i = lane; i < n; i += 32) put(max(src[i], 0));
        } else {
            const float loc = (float)a.kw.p1[pi], scale = (float)a.kw.p2[pi];
            const uint32_t c2 = stream_word(ST_IDEAL, 0u, (uint32_t)k);
            for (int blk = lane; 4 * blk < n; blk += 32) {
                const uint4 w = philox4x32_10((uint32_t)blk, a.step, c2, a.env_base + (uint32_t)e, k0, k1);
                put(laplace_cents(w.x, loc, scale));
                if (4 * blk + 1 < n) put(laplace_cents(w.y, loc, scale));
                if (4 * blk + 2 < n) put(laplace_cents(w.z, loc, scale));
                if (4 * blk + 3 < n) put(laplace_cents(w.w, loc, scale));
            }
        }
        __syncwarp();
        // ---- inclusive scan of (count, count x cents): every lane owns a run of 16 buckets
        constexpr int RUN = kIdealBuckets / 32;
        unsigned run_cnt = 0, run_sum = 0;
        for (int i = 0; i < RUN; ++i) {
            const int c = lane * RUN + i;
            run_cnt += cnt[c];
            run_sum += cnt[c] * (unsigned)c;
        }
        unsigned pre_cnt = run_cnt, pre_sum = run_sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned tc = __shfl_up_sync(FULL, pre_cnt, off), ts = __shfl_up_sync(FULL, pre_sum, off);
            if (lane >= off) { pre_cnt += tc; pre_sum += ts; }
        }
        unsigned acc_c = pre_cnt - run_cnt, acc_s = pre_sum - run_sum;
        for (int i = 0; i < RUN; ++i) {
            const int c = lane * RUN + i;
            const unsigned h = cnt[c];
            acc_c += h;
            acc_s += h * (unsigned)c;
            cnt[c] = acc_c;
            sum[c] = acc_s;
        }
        __syncwarp();
        const unsigned over_min = s_over_min[warp];
        // ---- the bid grid, 32 bids per trip
        const double vol = a.kw.vol_mean[pi], bctr = a.kw.ctr[pi], sctr = a.kw.cvr[pi], mrev = a.kw.rev_mean[pi];
        const double margin_rev = __dmul_rn(sctr, mrev);
        double best = -1.0;
        int best_g = 0, n_pos = 0;
        for (int g0 = 0; g0 < a.n_grid; g0 += 32) {
            const int g = g0 + lane;
            double prof = -1.0;
            if (g < a.n_grid) {
                const int cb = grid.thr[g];  // < kIdealBuckets - 1 (checked on the host)
                const unsigned idx = cb < 0 ? 0u : cnt[cb];           // searchsorted(side="right")
                unsigned tsum = cb < 0 ? 0u : sum[cb], take = idx;
                if (idx < (unsigned)n) {
                    // mean_prices[idx] averages idx + 1 sorted samples (metrics.py:31-35): the ones at
                    // or below the bid and the lowest one above it = the next non-empty bucket
                    int c = cb + 1;
                    while (c < kIdealBuckets - 1 && cnt[c] == idx) ++c;
                    tsum += c < kIdealBuckets - 1 ? (unsigned)c : over_min;
                    take = idx + 1;
                }
                const double rate = __ddiv_rn((double)idx, (double)n);
                const double cpc = __ddiv_rn(__ddiv_rn((double)tsum, 100.0), (double)take);
                // ((vol_mean * rate) * bctr) * (sctr * mean_rev - cpc), clipped at 0 (metrics.py:51-57)
                prof = __dmul_rn(__dmul_rn(__dmul_rn(vol, rate), bctr), __dsub_rn(margin_rev, cpc));
                if (!(prof > 0.0)) prof = 0.0;
                if (a.impression_rate) a.impression_rate[u * a.n_grid + g] = rate;
                if (a.expected_cpc) a.expected_cpc[u * a.n_grid + g] = cpc;
            }
            n_pos += __popc(__ballot_sync(FULL, prof > 0.0));
            double m = prof;  // warp arg-max, the first index wins ties like np.argmax
            int mg = g;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double om = __shfl_xor_sync(FULL, m, off);
                const int og = __shfl_xor_sync(FULL, mg, off);
                if (om > m || (om == m && og < mg)) { m = om; mg = og; }
            }
            if (m > best) { best = m; best_g = mg; }
        }
        if (lane == 0) {
            a.ideal_profit[u] = best > 0.0 ? best : 0.0;
            if (a.positive_frac) a.positive_frac[u] = __ddiv_rn((double)n_pos, (double)a.n_grid);
            if (a.best_bid_index) a.best_bid_index[u] = best_g;
        }
        __syncwarp();
    }
}

// Returns cudaErrorInvalidValue when a grid bid lies beyond the exact histogram range.
cudaError_t launch_ideal_profit(const adc_ideal_args &a, cudaStream_t s, int64_t *launches)
{
    IdealGrid grid;
    for (int g = 0; g < a.n_grid; ++g) {
        const double b = a.bid_grid_host[g];
        if (!(b == b) || b >= (kIdealBuckets - 2) / 100.0) return cudaErrorInvalidValue;
        long long c = (long long)floor(b * 100.0);
        while ((double)(c + 1) / 100.0 <= b) ++c;  // np.around(x, 2) values are c / 100 correctly rounded
        while (c >= 0 && (double)c / 100.0 > b) --c;
        grid.thr[g] = (short)(c < -1 ? -1 : c);
    }
    const int64_t total = (int64_t)a.E * a.kw.K;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adc_ideal_profit_kernel, kIdealWarps * 32, 0);
    if (per_sm < 1) per_sm = 1;
    int64_t nblk = (int64_t)sms * per_sm;
    const int64_t want = (total + kIdealWarps - 1) / kIdealWarps;
    if (want < nblk) nblk = want;
    if (nblk < 1) nblk = 1;
    adc_ideal_profit_kernel<<<(unsigned)nblk, kIdealWarps * 32, 0, s>>>(a, grid);
    ++*launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// AKNCP / NCP of a window from the episode accumulators (adc_episode_metrics): one warp per env.
// The K ratios go to shared memory; the two middle order statistics come from stable ranks
// (rank_i = #{j : r_j < r_i or (r_j == r_i and j < i)}: K^2 / 32 comparisons per warp, nothing is
// sorted or moved).  A CTA adds its envs' sums with six atomics.
// ------------------------------------------------------------------------------------------
constexpr int kMetWarps = 4;

__global__ void __launch_bounds__(kMetWarps * 32)
adc_episode_metrics_kernel(const __grid_constant__ adc_metrics_args a)
{
    extern __shared__ __align__(16) double m_ratio[];  // [kMetWarps][K]
    __shared__ double s_part[kMetWarps][6];
    const int K = a.K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *ratio = m_ratio + (size_t)warp * K;
    const unsigned FULL = 0xFFFFFFFFu;
    const double steps = (double)a.steps;
    double t_prof = 0.0, t_ideal = 0.0, t_ak = 0.0, t_ak2 = 0.0, t_ncp = 0.0, t_n = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * kMetWarps + warp; e < a.E; e += (int64_t)gridDim.x * kMetWarps) {
        int64_t *acc = a.episode_profit_cents + e * K;
        const double *idl = a.ideal + e * a.ideal_env_stride;
        double prof_s = 0.0, ideal_s = 0.0;
        for (int k = lane; k < K; k += 32) {
            const double prof = __ddiv_rn((double)acc[k], 100.0);
            const double id = idl[k];
            const double den = id <= 0.0 ? 1.0 : id;
            ratio[k] = __ddiv_rn(__ddiv_rn(prof, steps), den);
            prof_s += prof;
            ideal_s += id;
            if (a.zero) acc[k] = 0;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            prof_s += __shfl_xor_sync(FULL, prof_s, off);
            ideal_s += __shfl_xor_sync(FULL, ideal_s, off);
        }
        __syncwarp();
        // the two middle order statistics by stable rank
        const int r_lo = (K + 1) / 2 - 1, r_hi = K / 2;
        double lo = 0.0, hi = 0.0;
        for (int i = lane; i < K; i += 32) {
            const double x = ratio[i];
            int rank = 0;
            for (int j = 0; j < K; ++j) {
                const double y = ratio[j];
                rank += (y < x) || (y == x && j < i);
            }
            if (rank == r_lo) lo = x;
            if (rank == r_hi) hi = x;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {  // exactly one lane holds each of them: the others add 0.0
            lo += __shfl_xor_sync(FULL, lo, off);
            hi += __shfl_xor_sync(FULL, hi, off);
        }
        __syncwarp();
        const double akncp = 0.5 * (lo + hi);
        const double isum = ideal_s * steps;
        const double ncp = __ddiv_rn(prof_s, isum <= 0.0 ? 1.0 : isum);
        if (lane == 0) {
            if (a.akncp != nullptr) a.akncp[e] = akncp;
            if (a.ncp != nullptr) a.ncp[e] = ncp;
        }
        t_prof += prof_s; t_ideal += isum; t_ak += akncp; t_ak2 += akncp * akncp; t_ncp += ncp; t_n += 1.0;
    }
    if (lane == 0) {
        s_part[warp][0] = t_prof; s_part[warp][1] = t_ideal; s_part[warp][2] = t_ak;
        s_part[warp][3] = t_ak2; s_part[warp][4] = t_ncp; s_part[warp][5] = t_n;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kMetWarps; ++w) v += s_part[w][threadIdx.x];
        if (v != 0.0) atomicAdd(a.sums + threadIdx.x, v);
    }
}

cudaError_t launch_episode_metrics(const adc_metrics_args &a, cudaStream_t s, int64_t *launches)
{
    auto kern = adc_episode_metrics_kernel;
    const size_t dyn = (size_t)kMetWarps * (size_t)a.K * sizeof(double);
    static size_t configured = 0;
    if (dyn > 48 * 1024 && dyn > configured) {
        const cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (err != cudaSuccess) return err;
        configured = dyn;
    }
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t grid = ((int64_t)a.E + kMetWarps - 1) / kMetWarps;
    const int64_t cap = (int64_t)(sms > 0 ? sms : 148) * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kMetWarps * 32, dyn, s>>>(a);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace adc
