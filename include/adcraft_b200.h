/* adcraft_b200 -- C ABI of the B200-native BiddingSimulation.step hot path.
 *
 * This header is the drop-in boundary.  The reference (Mikata-Project/adcraft) crosses one
 * FFI on this path: the PyO3 module `adcraft.rust` (src/lib.rs:14-278), called per keyword and
 * per lane from the Python loops in adcraft/bidding_simulation.py:170-234.  On the GPU the whole
 * loop nest moves behind the boundary, so the entry points below replace, at batch
 * granularity (E envs x K keywords per call):
 *
 *   adc_step_philox   <- BiddingSimulation.step            adcraft/gymnasium_kw_env.py:160-269
 *                        simulate_epoch_of_bidding_on_campaign  adcraft/bidding_simulation.py:170-234
 *                        simulate_epoch_of_bidding              adcraft/bidding_simulation.py:44-120
 *                        ImplicitKeyword.auction / nth_price_auction
 *                                                     adcraft/synthetic_kw_classes.py:623-646,
 *                                                     adcraft/synthetic_kw_helpers.py:116-180
 *                        ExplicitKeyword.auction           adcraft/synthetic_kw_classes.py:493-538
 *                        rust.nonneg_int_normal_sampler    src/lib.rs:245-248,314-325
 *                        rust.binomial_impressions         src/lib.rs:69-76
 *                        rust.threshold_sigmoid            src/lib.rs:92-105
 *                        rust.cost_create                  src/lib.rs:53-67
 *                        rust.sum_list/sum_array/sum_array_bool  src/lib.rs:107-127
 *                        update_keywords (drift)           adcraft/gymnasium_kw_env.py:114-158
 *   adc_step_replay   <- the same path, fed pre-drawn volumes / competitor bids / uniforms /
 *                        revenues (parity mode; the reference's RNG is unseedable, src/lib.rs:316-320)
 *   adc_reset_envs    <- BiddingSimulation.reset (state part) adcraft/gymnasium_kw_env.py:326-329
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ or torch types.
 *   - every array pointer is a CALLER-OWNED DEVICE pointer unless its name ends in _host;
 *     the library allocates nothing and keeps no state between calls.
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*); they return after
 *     enqueueing.  Return value: 0 on success, a negative adc_status otherwise, with a
 *     thread-local message available from adc_last_error().  Nothing throws across the ABI.
 *   - money is carried in integer cents inside the kernels (competitor bids, costs and revenues
 *     are cent-rounded in the reference: synthetic_kw_helpers.py:68-70,108-113); float outputs
 *     are cents/100 converted once.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     ADC_ERR_NO_DEVICE.
 */
#ifndef ADCRAFT_B200_H
#define ADCRAFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADC_ABI_VERSION 6
#define ADC_SUBSTEPS 24 /* adcraft/bidding_simulation.py:213 */

typedef enum adc_status {
    ADC_OK = 0,
    ADC_ERR_INVALID = -1,   /* bad argument (null pointer, size, enum)       */
    ADC_ERR_NO_DEVICE = -2, /* no CUDA device / driver                        */
    ADC_ERR_CUDA = -3,      /* a CUDA call or launch failed                   */
    ADC_ERR_UNSUPPORTED = -4
} adc_status;

/* keyword kind (synthetic_kw_classes.py:457,578).  ADC_IMPLICIT is the experiments' ImplicitKeyword
 * (one competitor, |Laplace| rounded to cents: gymnasium_kw_utils.py:159-166, helpers:104-113);
 * ADC_IMPLICIT_MULTI is the class default (classes:649-688): m ~ Binomial(max_bidders,
 * participation) bidders drawn once per (sub-step, keyword) lane, signed un-rounded
 * Laplace(loc, scale) bids, cleared by nth_price_auction(n=2, num_winners=1) incl. its zero
 * padding for m < 3 (helpers:156-161). */
enum { ADC_IMPLICIT = 0, ADC_EXPLICIT = 1, ADC_IMPLICIT_MULTI = 2 };
enum { ADC_F32 = 0, ADC_F64 = 1 };            /* dtype tags for bids / float outputs             */

/* Keyword parameters, SoA float64 (gymnasium_kw_utils.py:20-28).  env_stride = 0: one keyword set
 * [K] shared by every env; env_stride = K: per-env sets [E,K] (required when drift is on).
 * vol_mean, ctr, cvr are updated in place by the drift. */
typedef struct adc_keywords {
    int32_t kind;
    int32_t K;
    int64_t env_stride;
    double *vol_mean;
    const double *vol_std;   /* also the drift scale "init_volumes" (gymnasium_kw_env.py:136-137) */
    const double *p1;        /* implicit: Laplace loc   | explicit: impression_bid_intercept */
    const double *p2;        /* implicit: Laplace scale | explicit: impression_slope         */
    double *ctr;
    double *cvr;
    const double *rev_mean;
    const double *rev_std;
    const double *max_bidders;   /* ADC_IMPLICIT_MULTI only, else NULL (classes:659-662, default 30)  */
    const double *participation; /* ADC_IMPLICIT_MULTI only, else NULL (classes:663, default 3/5)     */
    double impression_thresh; /* explicit only (0.05, gymnasium_kw_utils.py:81) */
} adc_keywords;

/* Per-env state [E] (gymnasium_kw_env.py:80,327-328). */
typedef struct adc_env_state {
    double *budget;      /* persisted budget ("this timestep and onward", env:197-199) */
    double *cum_profit;
    int32_t *day;
    int32_t max_days;
    double loss_threshold;
} adc_env_state;

/* Drift (update_keywords, env:114-158).  mask == NULL: stationary. */
typedef struct adc_drift {
    const uint8_t *mask;  /* [K] */
    int32_t num_updates;  /* sum(mask): zip() truncation, only keywords < num_updates considered */
    double mag[3];        /* updater_params: vol, ctr, cvr half-widths */
} adc_drift;

/* Outputs of one step.  [E,K] unless noted; optional pointers may be NULL. */
typedef struct adc_step_out {
    int32_t *impressions;
    int32_t *clicks;        /* buyside_clicks        */
    int32_t *conversions;   /* sellside_conversions  */
    void *cost;             /* float_dtype           */
    void *revenue;          /* float_dtype           */
    int32_t float_dtype;    /* ADC_F32 / ADC_F64     */
    int64_t *cost_cents;    /* REQUIRED exact cents accumulator (explicit keywords: unused, 0) */
    int64_t *revenue_cents; /* REQUIRED exact cents accumulator                                */
    double *reward;         /* [E] */
    double *obs_cum_profit; /* [E] value reported in the observation (before auto-reset) */
    int32_t *obs_days;      /* [E] */
    uint8_t *terminated;    /* [E] */
    uint8_t *truncated;     /* [E] */
    double *remaining_budget; /* [E] optional: budget left after the day */
    int64_t *episode_profit_cents; /* [E,K] optional running sum: += revenue_cents - cost_cents of every
                                 step, added when the env's step is final (the per-keyword profits
                                 AKNCP / NCP are made of, experiment_metrics.py:64-83); the caller
                                 zeroes it at the episode boundaries it cares about */
    double *episode_reward;   /* [E] optional running sum of the env's rewards (added in the env tail)   */
    int32_t *episode_count;   /* [E] optional running count of finished episodes (terminated | truncated) */
    void *rows;               /* optional [E, adc_host_row_bytes] compact observation rows (the layout of
                                 adc_host_chunk below), packed by the warp that finalises the env.  May be
                                 HOST memory mapped into the device address space (pinned, UVA): the step
                                 then delivers its observations to the host while it runs, 14 bytes per
                                 unit instead of 20, and no copy is enqueued */
    void *flat_obs;           /* optional [E, 5K+2] float_dtype: the reference's flat observation row
                                 (FlatArrayWrapper, wrappers/flat_array.py:44-87; keys sorted like
                                 gymnasium_kw_utils.py:383-390): buyside_clicks[0:K] | cost[K:2K] |
                                 cumulative_profit | days_passed | impressions | revenue |
                                 sellside_conversions -- written by the kernels, no gather pass */
    void *unit_records;       /* optional [E,K] adc_unit_record (16 bytes per unit): the unit's whole
                                 observation in ONE aligned 16-byte store, 512 contiguous bytes per warp.
                                 Meant to be HOST memory mapped into the device address space (pinned,
                                 UVA): with reward / obs_cum_profit / obs_days / terminated / truncated
                                 pointing at host memory as well, the step delivers its observations over
                                 PCIe while it runs -- 16 bytes per unit in full-size write transactions
                                 instead of 20 in five 4-byte streams -- and the int32 / float arrays above
                                 stay in device memory.  Needs float_dtype ADC_F32 */
} adc_step_out;

/* One unit of adc_step_out.unit_records (little endian).  A count above 65535 is stored as 65535 with
 * bit 0 of `flags` set (the device arrays keep the exact int32 values). */
typedef struct adc_unit_record {
    uint16_t impressions, clicks, conversions, flags;
    float cost, revenue;
} adc_unit_record;

/* Scratch the step needs (caller-owned so that nothing is allocated per call). */
typedef struct adc_scratch {
    int32_t *serial_list;   /* [E] envs that need the exact serial budget walk              */
    int32_t *serial_count;  /* [2] double-buffered on step parity; both 0 before the first step */
    int64_t *env_profit;    /* [E] per-env profit cents accumulator, 0 on entry and on exit */
    int64_t *env_cost;      /* [E] per-env cost cents accumulator,   0 on entry and on exit */
    int32_t *env_done;      /* [E] finished-unit counter,            0 on entry and on exit */
    double *unit_cost_f64;  /* [E,K] explicit keywords only: un-rounded cost sums (else NULL) */
    /* Optional (all three or none), [E,K] DEVICE memory: running impressions / clicks / conversions
     * of the exact serial walk.  NULL: the walk read-modify-writes out.impressions / clicks /
     * conversions themselves, which is fine for device outputs; with outputs in mapped host memory
     * every sub-step would cross PCIe twice, so give the walk device scratch here and it stores the
     * day's totals to the outputs once per env. */
    /* Optional [2] DEVICE words, both 0 before the first step (double-buffered on the step parity
     * like serial_count): with it the hot kernel's warps pull their batches from an atomic
     * counter instead of a static round-robin deal (results are identical either way). */
    uint32_t *work_counter;
    int32_t *acc_impressions;
    int32_t *acc_clicks;
    int32_t *acc_conversions;
    /* Optional DEVICE workspace of the warp-cooperative exact serial walk (free-running implicit
     * keywords): every resident warp expands one queued env's day into a slab; the smallest
     * slab is adc_serial_slab_bytes(K) (about 630 bytes per keyword, 128 clicked slots per keyword), and
     * the warps that run share the whole workspace: room beyond the minimum goes to the slabs' slot pools
     * (up to 512 slots per keyword, where no day of <= 512 auctions can overflow them).  16-byte aligned.
     * The launcher runs as many warps as smallest slabs fit (about 4100 resident warps at most); NULL / too small for one slab: the walk falls back to one
     * thread per env (correct, much slower). */
    void *serial_ws;
    int64_t serial_ws_bytes;
    /* Optional [E] bytes, zeroed once by the caller and then left to the library (free-running implicit
     * keywords with serial_ws): the exact walk marks the envs whose budget really bound; in the next step
     * the budget-free kernel does not evaluate them at all and queues them for the walk right away.  Both
     * kernels compute the same function, so results do not depend on the marks -- only on how often a
     * budget-bound env is evaluated twice. */
    uint8_t *serial_hint;
    /* Optional [E,K] DEVICE bytes, used with env_group > 1 (shared auctions, free-running, no floor_cents
     * table): a compact pre-pass finds every (world, keyword)'s unique top bidder, finishes the units of
     * all the others (zero outputs, env completion) and marks them here; the hot kernel skips marked
     * units.  Without it every unit goes through the hot kernel, whose instruction caches thrash when
     * seven of eight batches only run its straight-line setup and output code (C4: 4.6 ms). */
    uint8_t *outbid_mask;
} adc_scratch;

/* Optional per-click detail (the ragged lists of BiddingOutcomes, bidding_simulation.py:10-38, that
 * the reference prints in info["bidding_outcomes"]).  Only the exact serial path fills it: call
 * with force_serial = 1 and n_lanes = 1.  All pointers NULL: not recorded. */
typedef struct adc_detail {
    int32_t cap;            /* clicks recorded per (env, keyword) and step, at most          */
    double *costs;          /* [E,K,cap] cost of each accepted click, in order (bsim:101)    */
    double *rev_per_cost;   /* [E,K,cap] revenue of that click, 0 if it did not convert (:114-115) */
    int32_t *n_recorded;    /* [E,K] min(clicks, cap)                                        */
    double *volume_seen;    /* [E,K] combine_outcomes' share denominator (bsim:130-146): the
                               auctions of the lanes that had at least one impression        */
    int32_t *lane_clicks;   /* [E,K,24] accepted clicks of each (sub-step) lane, 0 for lanes after
                               the early break: with lane_convs it splits the two lists per lane,
                               which is what the reference's per-lane profit sums need (bsim:117,138) */
    int32_t *lane_convs;    /* [E,K,24] conversions of each lane                                */
} adc_detail;

typedef struct adc_step_args {
    int32_t E;              /* envs owned by this call / rank                      */
    uint32_t env_base;      /* global id of env 0 (Philox counter; rank sharding)  */
    uint32_t step;          /* Philox counter word: steps since the last seeded reset, so that
                               reset(seed=s) replays the same trajectory                       */
    uint32_t parity;        /* must flip (advance by 1) on every call that uses a given scratch:
                               its low bit double-buffers serial_count / work_counter          */
    int32_t device;         /* CUDA device ordinal every pointer lives on; the call fails with
                               ADC_ERR_INVALID when it is not the calling thread's current
                               device.  -1: not checked                                        */
    uint64_t seed;          /* Philox key                                          */
    int32_t n_lanes;        /* 0: default.  1: queued envs take the one-thread-per-env exact walk (the only
                               one that records adc_detail).  Other values of the round-1 ABI (-8, -16,
                               -32, 2..32) are accepted and mean 0 */
    int32_t budget_alias;   /* 1: ndarray-budget double charge (bsim:102 + :225), 0: scalar budget */
    int32_t autoreset;      /* 1: zero cum_profit/day of finished envs after reporting them */
    int32_t force_serial;   /* 1: run every env through the exact serial kernel (testing)   */
    adc_keywords kw;
    adc_env_state env;
    adc_drift drift;
    const void *bids;       /* [E,K] dollars, canonicalised to cents inside (env:215) */
    int32_t bids_dtype;     /* ADC_F32 / ADC_F64 */
    int32_t f32_ties;       /* 1 (with ADC_F32 bids): numpy >= 2 keeps a float32 bid float32 through
                               np.maximum / round (env:215), so searchsorted compares float32(cents/100)
                               upcast to float64 with the float64 competitor bids: a bid whose float32
                               value lies above its cent value WINS ties (0.30f > 0.30), one below
                               loses them.  0: float64 semantics, ties always lose */
    const void *budget_in;  /* optional [E] dollars, same dtype as bids: rounded to cents and stored */
    /* Shared auctions (several bidders in ONE auction; free-running implicit keywords only).
     * env_group = A > 1: envs [g*A, (g+1)*A) are the A bidders of world g and share every draw
     * (Philox env id = env_base + e / A: same volumes, same competitor bids, one click/conversion
     * word per auction, consumed by the single winner).  floor_cents[e,k] = the highest rival bid in
     * cents: the clearing price of an auction is max(sampled competitor, floor), i.e.
     * nth_price_auction(n=2, num_winners=1) on rivals + competitor (synthetic_kw_helpers.py:116-180).
     * 0 / NULL: independent envs. */
    int32_t env_group;
    int32_t spread_outcomes; /* free-running implicit keywords: 1 = a batch whose days are uneven (dense keywords,
                                volumes 128 +- 60: the longest day has 8 groups of 32 auctions where the average
                                has 4.5) spreads its (unit, group) pairs over the lanes instead of running the
                                longest day's group count for all 32 units; same draws, same results, 7 % faster
                                on such sets and 2.4 % slower on sparse ones (which never spread): 0 for those */
    const int32_t *floor_cents;
    adc_step_out out;
    adc_scratch scratch;
    adc_detail detail;
} adc_step_args;

/* Replay tape for E envs, consumption order, CSR over the E*K units (u = e*K + k):
 * stream[off[u] .. off[u+1]).  Streams may be longer than what gets consumed.
 *
 * Optional packed copy (implicit keywords): the same values laid out as ONE record per unit, so a
 * unit's whole day is one contiguous, 16-byte aligned block that the replay kernel fetches with a
 * single bulk copy (TMA) into shared memory.  Record of unit u = packed[packed_off[u] ..
 * packed_off[u+1]) (byte offsets, multiples of 16; an empty record means volume 0):
 *     int32  hdr[8]    = { V, n_comp, n_click, n_conv, n_rev, flags, 0, 0 }   with n_comp <= V;
 *                        flags bit 0 (ADC_PACKED_NARROW): no competitor bid of the record is negative and
 *                        every revenue lies in [0, 65535] cents -- the replay kernel then walks the record
 *                        with 32-bit sums without looking for such values; a record without the flag takes
 *                        the generic 64-bit walk (same results)
 *     int32  comp[n_comp rounded up to a multiple of 4]   padding entries = INT32_MAX
 *     double click[n_click];  double conv[n_conv]
 *     int32  rev[n_rev]       zero padded to the next multiple of 16 bytes
 * `packed` must be 16-byte aligned.  The CSR streams stay mandatory: envs whose budget may bind and
 * records that fail validation are re-walked by the exact serial kernel from the CSR form. */
#define ADC_PACKED_NARROW 1
typedef struct adc_tape {
    const int32_t *volume;                              /* [E,K]                          */
    const int64_t *comp_off;  const int32_t *comp_cents;   /* implicit: one per auction      */
    const double *comp_f64;   /* ADC_IMPLICIT_MULTI (shares comp_off): the auction's clearing price,
                                 max(other bids, and 0 when fewer than 3 bidders) -- helpers:156-177 */
    const int64_t *click_off; const double *u_click;       /* one per click slot             */
    const int64_t *conv_off;  const double *u_conv;        /* one per accepted click         */
    const int64_t *rev_off;   const int32_t *rev_cents;    /* one per conversion             */
    const int32_t *impr;                                /* explicit: [E,K,24] impressions */
    const int64_t *cost_off;  const double *cost;          /* explicit: one per impression   */
    const double *drift;                                /* optional [E,3,K] coefficients  */
    const unsigned char *packed;                        /* optional packed records, or NULL */
    const int64_t *packed_off;                          /* [E*K+1] byte offsets into packed */
} adc_tape;

/* Ideal-profit estimator behind the AKNCP / NCP metrics (experiment_metrics.py:20-61): per (env,
 * implicit keyword) n_samples competitor bids -> sort -> searchsorted(side="right") of every grid bid
 * -> impression rate and mean price of the idx + 1 lowest samples -> expected profit
 * vol_mean * rate * bctr * (sctr * mean_rev - cpc) clipped at 0 -> max over the grid.  Recomputed
 * per step when the keywords drift (the reference's notebooks call it once per reset). */
#define ADC_IDEAL_MAX_GRID 512
typedef struct adc_ideal_args {
    int32_t E;
    uint32_t env_base;
    uint32_t step;               /* Philox counter word of the sampling stream                  */
    int32_t device;              /* like adc_step_args.device                                   */
    uint64_t seed;
    adc_keywords kw;             /* ADC_IMPLICIT; reads p1, p2, vol_mean, ctr, cvr, rev_mean      */
    int32_t n_samples;           /* 2048 in the reference (metrics.py:21), <= 65536             */
    int32_t n_grid;              /* <= ADC_IDEAL_MAX_GRID                                       */
    const double *bid_grid_host; /* [n_grid] HOST array, e.g. np.arange(0.01, 3.00, 0.01); every bid
                                    below 5.10 (the exact range of the counting sort)           */
    const int32_t *samples_cents; /* optional [E,K,n_samples] pre-drawn competitor bids in cents
                                     (parity mode); NULL: drawn from the Philox stream         */
    double *ideal_profit;        /* [E,K] max over the grid, >= 0 (metrics.py:58)               */
    double *positive_frac;       /* optional [E,K] share of grid bids with positive profit (:59) */
    int32_t *best_bid_index;     /* optional [E,K] argmax (:59)                                  */
    double *impression_rate;     /* optional [E,K,n_grid] (metrics.py:31)                        */
    double *expected_cpc;        /* optional [E,K,n_grid] (metrics.py:34-35)                     */
} adc_ideal_args;

const char *adc_last_error(void);
int adc_abi_version(void);
int adc_device_count(void);
/* sizeof(adc_step_args) / sizeof(adc_tape) as compiled: lets an FFI binding check its layout. */
int adc_sizeof_step_args(void);
int adc_sizeof_tape(void);
int adc_sizeof_ideal_args(void);
/* Bytes of adc_scratch.serial_ws one slab (one resident warp of the exact serial walk) needs for K keywords. */
int64_t adc_serial_slab_bytes(int32_t K);

/* One free-running env step for E envs (counter-based Philox draws keyed by
 * (seed, env_base+e, keyword, step)).  Launches: fused lane kernel, then the exact serial kernel
 * for envs whose budget may bind (skips itself when the list is empty). */
int adc_step_philox(const adc_step_args *args, void *stream);

/* The same step driven by a pre-drawn tape (parity mode). seed/step in args are ignored. */
int adc_step_replay(const adc_step_args *args, const adc_tape *tape, void *stream);

/* Reset per-env episode state of the envs with mask[e] != 0 (mask NULL: all). */
int adc_reset_envs(int32_t E, const uint8_t *mask, double *cum_profit, int32_t *day, void *stream);

/* ---- host round trip (the call a CPU-side RL loop makes) --------------------------------------
 * adc_step_host runs one free-running step whose bids come from HOST memory and whose observations
 * land in HOST memory, as a pipeline over chunks of envs: for every chunk, on the chunk's stream,
 *   cudaMemcpyAsync(bids: host -> device)  ->  the step kernels  ->  adc_pack_rows  ->
 *   cudaMemcpyAsync(rows: device -> host)
 * so the copy engines move chunk i while the SMs work on chunk i + 1 (PCIe is full duplex: the two
 * directions overlap too).  One cudaMemcpyAsync per chunk and direction.  Returns after every
 * chunk's rows have landed (it synchronises the chunk streams).
 *
 * Compact row of one env (row_bytes = adc_host_row_bytes(K, float_dtype); little endian):
 *     uint16 impressions[K] | uint16 clicks[K] | uint16 conversions[K] | pad to 8 bytes
 *     float  cost[K] | float revenue[K]          (float_dtype ADC_F32; doubles with ADC_F64) | pad to 8 bytes
 *     double reward | double cumulative_profit | int32 days_passed
 *     uint8 terminated | uint8 truncated | uint8 count_overflow | uint8 0
 * count_overflow = 1 when a count of the env exceeds 65535 (the uint16 then holds 65535; the device
 * arrays of adc_step_out keep the exact int32 values). */
typedef struct adc_host_chunk {
    adc_step_args args;       /* the chunk's envs: every pointer already offset to its first env, its own
                                 scratch counters (serial_count, work_counter); args.bids = DEVICE staging */
    const void *bids_host;    /* [args.E, K] pinned host bids, args.bids_dtype */
    void *rows_dev;           /* [args.E, row_bytes] device staging */
    void *rows_host;          /* [args.E, row_bytes] pinned host output */
    void *stream;             /* the chunk's stream (chunks may share streams; >= 2 distinct ones overlap) */
} adc_host_chunk;

int64_t adc_host_row_bytes(int32_t K, int32_t float_dtype);
int adc_step_host(const adc_host_chunk *chunks, int32_t n_chunks);
int adc_sizeof_host_chunk(void);

/* The ideal-profit estimator above for E x K units. */
int adc_ideal_profit(const adc_ideal_args *args, void *stream);

/* AKNCP / NCP of a window of steps (experiment_metrics.py:64-83) from the step kernels' per-keyword
 * episode accumulators: per env  AKNCP = median over its keywords of (mean profit / mean ideal profit,
 * an ideal <= 0 counts as 1)  and  NCP = sum profit / sum ideal (a sum <= 0 counts as 1); the median is
 * np.median's (mean of the two middle values for an even K), taken per env, never across ranks.
 * One warp per env; the six sums a rank contributes to the NCCL all-reduce are ADDED to `sums`:
 *   sums[0] += sum profit, [1] += sum ideal * steps, [2] += sum AKNCP_e, [3] += sum AKNCP_e^2,
 *   [4] += sum NCP_e, [5] += E                 (dollars; entries 6 and 7 belong to the caller). */
#define ADC_METRICS_MAX_K 2048
typedef struct adc_metrics_args {
    int32_t E, K;                        /* K <= ADC_METRICS_MAX_K                                        */
    int32_t steps;                       /* steps accumulated in episode_profit_cents (> 0)              */
    int32_t device;                      /* like adc_step_args.device                                    */
    int64_t *episode_profit_cents;       /* [E,K] adc_step_out.episode_profit_cents                      */
    const double *ideal;                 /* per-step ideal profit, [K] (ideal_env_stride 0) or [E,K] (K) */
    int64_t ideal_env_stride;
    double *sums;                        /* [8] device doubles, see above                                */
    double *akncp;                       /* optional [E] per-env AKNCP                                   */
    double *ncp;                         /* optional [E] per-env NCP                                     */
    int32_t zero;                        /* != 0: zero episode_profit_cents after reading (next window)  */
    int32_t pad_;
} adc_metrics_args;
int adc_episode_metrics(const adc_metrics_args *args, void *stream);
int adc_sizeof_metrics_args(void);

/* Number of kernel launches issued by this library on the calling thread since the last call
 * with reset != 0 (bench.py's gpu_launches). */
int64_t adc_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* ADCRAFT_B200_H */
