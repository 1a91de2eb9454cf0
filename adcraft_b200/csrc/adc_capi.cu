// extern "C" entry points declared in include/adcraft_b200.h.
// Argument validation + error reporting live here; kernels and launch logic in adc_step.cu.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "adc_step.h"

namespace {

thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;

int fail(adc_status st, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return (int)st;
}

int check_device()
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        cudaGetLastError();
        return fail(ADC_ERR_NO_DEVICE, "adcraft_b200: no CUDA device (%s); there is no CPU fallback",
                    e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    }
    return ADC_OK;
}

#define ADC_REQUIRE(cond, what)                                                                   \
    do {                                                                                          \
        if (!(cond)) return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: %s", what);    \
    } while (0)

int validate(const adc_step_args *a, const adc_tape *tape)
{
    ADC_REQUIRE(a != nullptr, "args is NULL");
    ADC_REQUIRE(a->E > 0, "E must be > 0");
    ADC_REQUIRE(a->kw.K > 0 && a->kw.K < (1 << 20), "K must be in [1, 2^20)");
    ADC_REQUIRE(a->kw.kind == ADC_IMPLICIT || a->kw.kind == ADC_EXPLICIT || a->kw.kind == ADC_IMPLICIT_MULTI,
                "kw.kind");
    ADC_REQUIRE(a->kw.kind != ADC_IMPLICIT_MULTI || (a->kw.max_bidders && a->kw.participation),
                "kw.max_bidders / kw.participation are required for ADC_IMPLICIT_MULTI");
    ADC_REQUIRE(a->f32_ties == 0 || a->floor_cents == nullptr, "f32_ties is not defined for shared auctions");
    ADC_REQUIRE(a->kw.env_stride == 0 || a->kw.env_stride == a->kw.K, "kw.env_stride must be 0 or K");
    ADC_REQUIRE(a->kw.vol_mean && a->kw.vol_std && a->kw.p1 && a->kw.p2 && a->kw.ctr && a->kw.cvr &&
                    a->kw.rev_mean && a->kw.rev_std, "kw parameter pointer is NULL");
    ADC_REQUIRE(a->env.budget && a->env.cum_profit && a->env.day, "env state pointer is NULL");
    ADC_REQUIRE(a->bids != nullptr, "bids is NULL");
    ADC_REQUIRE(a->bids_dtype == ADC_F32 || a->bids_dtype == ADC_F64, "bids_dtype");
    ADC_REQUIRE(a->out.float_dtype == ADC_F32 || a->out.float_dtype == ADC_F64, "out.float_dtype");
    ADC_REQUIRE(a->out.impressions && a->out.clicks && a->out.conversions && a->out.cost &&
                    a->out.revenue && a->out.cost_cents && a->out.revenue_cents && a->out.reward &&
                    a->out.obs_cum_profit && a->out.obs_days && a->out.terminated && a->out.truncated,
                "output pointer is NULL");
    ADC_REQUIRE(a->out.unit_records == nullptr ||
                    (a->out.float_dtype == ADC_F32 && ((uintptr_t)a->out.unit_records & 15) == 0),
                "out.unit_records needs float_dtype ADC_F32 and 16-byte alignment");
    ADC_REQUIRE(a->scratch.serial_list && a->scratch.serial_count && a->scratch.env_profit &&
                    a->scratch.env_cost && a->scratch.env_done, "scratch pointer is NULL");
    ADC_REQUIRE(a->kw.kind == ADC_IMPLICIT || a->scratch.unit_cost_f64 != nullptr,
                "scratch.unit_cost_f64 is required for explicit and multi-bidder keywords (un-rounded costs)");
    ADC_REQUIRE((a->scratch.acc_impressions != nullptr) == (a->scratch.acc_clicks != nullptr) &&
                    (a->scratch.acc_clicks != nullptr) == (a->scratch.acc_conversions != nullptr),
                "scratch.acc_impressions / acc_clicks / acc_conversions: give all three or none");
    ADC_REQUIRE(a->scratch.serial_ws == nullptr ||
                    ((reinterpret_cast<uintptr_t>(a->scratch.serial_ws) & 15u) == 0 && a->scratch.serial_ws_bytes >= 0),
                "scratch.serial_ws must be 16-byte aligned");
    ADC_REQUIRE(a->drift.mask == nullptr || a->kw.env_stride == a->kw.K,
                "drift needs per-env keyword parameters (kw.env_stride == K)");
    ADC_REQUIRE(a->drift.mask == nullptr || (a->drift.num_updates >= 0 && a->drift.num_updates <= a->kw.K),
                "drift.num_updates");
    if (a->n_lanes > 0) {
        const int L = a->n_lanes;
        ADC_REQUIRE(L >= 1 && L <= 32 && (L & (L - 1)) == 0, "n_lanes must be a power of two <= 32");
    } else {
        ADC_REQUIRE(a->n_lanes == 0 || a->n_lanes == -8 || a->n_lanes == -16 || a->n_lanes == -32,
                    "n_lanes <= 0 selects the batched kernel: 0, -8, -16 or -32");
    }
    if (a->detail.costs != nullptr)
        ADC_REQUIRE(a->detail.cap > 0 && a->detail.rev_per_cost && a->detail.n_recorded && a->detail.volume_seen &&
                        a->detail.lane_clicks && a->detail.lane_convs, "detail: give every array or none");
    ADC_REQUIRE(a->env_group >= 0, "env_group must be >= 0");
    ADC_REQUIRE(a->env_group <= 1 || a->E % a->env_group == 0, "E must be a multiple of env_group");
    ADC_REQUIRE(a->f32_ties == 0 || a->env_group <= 1, "f32_ties is not defined for shared auctions");
    if (a->env_group > 1 || a->floor_cents != nullptr)
        ADC_REQUIRE(tape == nullptr && a->kw.kind == ADC_IMPLICIT,
                    "shared auctions (env_group / floor_cents) need free-running implicit keywords");
    if (tape) {
        ADC_REQUIRE(tape->volume && tape->click_off && tape->u_click && tape->conv_off && tape->u_conv &&
                        tape->rev_off && tape->rev_cents, "tape stream pointer is NULL");
        if (a->kw.kind == ADC_IMPLICIT)
            ADC_REQUIRE(tape->comp_off && tape->comp_cents, "tape.comp_* required for implicit keywords");
        else if (a->kw.kind == ADC_IMPLICIT_MULTI)
            ADC_REQUIRE(tape->comp_off && tape->comp_f64 && tape->impr,
                        "tape.comp_off / comp_f64 / impr (bidders per lane) required for multi-bidder keywords");
        else
            ADC_REQUIRE(tape->impr && tape->cost_off && tape->cost, "tape.impr/cost_* required for explicit keywords");
        ADC_REQUIRE(a->drift.mask == nullptr || tape->drift != nullptr, "tape.drift required when drift is on");
        if (tape->packed != nullptr) {
            ADC_REQUIRE(tape->packed_off != nullptr, "tape.packed_off is NULL");
            ADC_REQUIRE((reinterpret_cast<uintptr_t>(tape->packed) & 15u) == 0, "tape.packed must be 16-byte aligned");
            ADC_REQUIRE(a->kw.kind == ADC_IMPLICIT, "tape.packed is for implicit keywords");
        }
    }
    return ADC_OK;
}

int run(const adc_step_args *args, const adc_tape *tape, void *stream)
{
    int rc = validate(args, tape);
    if (rc) return rc;
    rc = check_device();
    if (rc) return rc;
    if (args->device >= 0) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != args->device)
            return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: the buffers live on device %d but the "
                        "calling thread's current device is %d", args->device, cur);
    }
    const cudaError_t e = adc::launch_step(*args, tape, static_cast<cudaStream_t>(stream), &g_launches);
    if (e != cudaSuccess) return fail(ADC_ERR_CUDA, "adcraft_b200: launch failed: %s", cudaGetErrorString(e));
    return ADC_OK;
}

}  // namespace

extern "C" {

const char *adc_last_error(void) { return g_err; }

int adc_abi_version(void) { return ADC_ABI_VERSION; }

int adc_sizeof_step_args(void) { return (int)sizeof(adc_step_args); }

int adc_sizeof_tape(void) { return (int)sizeof(adc_tape); }

int adc_sizeof_ideal_args(void) { return (int)sizeof(adc_ideal_args); }

int64_t adc_serial_slab_bytes(int32_t K) { return K > 0 ? adc::serial_slab_bytes(K) : 0; }

int adc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int adc_step_philox(const adc_step_args *args, void *stream) { return run(args, nullptr, stream); }

int adc_step_replay(const adc_step_args *args, const adc_tape *tape, void *stream)
{
    if (tape == nullptr) return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: tape is NULL");
    return run(args, tape, stream);
}

int adc_reset_envs(int32_t E, const uint8_t *mask, double *cum_profit, int32_t *day, void *stream)
{
    if (E <= 0 || !cum_profit || !day) return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: reset_envs");
    const int rc = check_device();
    if (rc) return rc;
    const cudaError_t e =
        adc::launch_reset_envs(E, mask, cum_profit, day, static_cast<cudaStream_t>(stream), &g_launches);
    if (e != cudaSuccess) return fail(ADC_ERR_CUDA, "adcraft_b200: launch failed: %s", cudaGetErrorString(e));
    return ADC_OK;
}

int64_t adc_host_row_bytes(int32_t K, int32_t float_dtype) { return K > 0 ? adc::host_row_bytes(K, float_dtype) : 0; }

int adc_sizeof_host_chunk(void) { return (int)sizeof(adc_host_chunk); }

int adc_step_host(const adc_host_chunk *chunks, int32_t n_chunks)
{
    if (chunks == nullptr || n_chunks <= 0)
        return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: adc_step_host needs at least one chunk");
    for (int i = 0; i < n_chunks; ++i) {
        const adc_host_chunk &c = chunks[i];
        const int rc = validate(&c.args, nullptr);
        if (rc) return rc;
        ADC_REQUIRE(c.bids_host && c.rows_dev && c.rows_host, "adc_step_host: chunk buffer is NULL");
        ADC_REQUIRE(c.args.kw.kind == ADC_IMPLICIT || c.args.kw.kind == ADC_EXPLICIT || c.args.kw.kind == ADC_IMPLICIT_MULTI,
                    "kw.kind");
    }
    int rc = check_device();
    if (rc) return rc;
    if (chunks[0].args.device >= 0) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != chunks[0].args.device)
            return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: the buffers live on device %d but the "
                        "calling thread's current device is %d", chunks[0].args.device, cur);
    }
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n_chunks && e == cudaSuccess; ++i) {
        const adc_host_chunk &c = chunks[i];
        cudaStream_t st = static_cast<cudaStream_t>(c.stream);
        const size_t bid_bytes = (size_t)c.args.E * c.args.kw.K * (c.args.bids_dtype == ADC_F64 ? 8 : 4);
        const size_t row_bytes = (size_t)c.args.E * (size_t)adc::host_row_bytes(c.args.kw.K, c.args.out.float_dtype);
        e = cudaMemcpyAsync(const_cast<void *>(c.args.bids), c.bids_host, bid_bytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = adc::launch_step(c.args, nullptr, st, &g_launches);
        if (e == cudaSuccess) e = adc::launch_pack_rows(c.args, c.rows_dev, st, &g_launches);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c.rows_host, c.rows_dev, row_bytes, cudaMemcpyDeviceToHost, st);
    }
    for (int i = 0; i < n_chunks; ++i) {  // the call returns host data: wait for every chunk's stream
        const cudaError_t es = cudaStreamSynchronize(static_cast<cudaStream_t>(chunks[i].stream));
        if (e == cudaSuccess) e = es;
    }
    if (e != cudaSuccess) return fail(ADC_ERR_CUDA, "adcraft_b200: adc_step_host failed: %s", cudaGetErrorString(e));
    return ADC_OK;
}

int adc_ideal_profit(const adc_ideal_args *a, void *stream)
{
    ADC_REQUIRE(a != nullptr, "args is NULL");
    ADC_REQUIRE(a->E > 0 && a->kw.K > 0, "E and K must be > 0");
    ADC_REQUIRE(a->kw.kind == ADC_IMPLICIT, "the ideal-profit estimator is defined for ADC_IMPLICIT keywords");
    ADC_REQUIRE(a->kw.env_stride == 0 || a->kw.env_stride == a->kw.K, "kw.env_stride must be 0 or K");
    ADC_REQUIRE(a->kw.vol_mean && a->kw.p1 && a->kw.p2 && a->kw.ctr && a->kw.cvr && a->kw.rev_mean,
                "kw parameter pointer is NULL");
    ADC_REQUIRE(a->n_samples > 0 && a->n_samples <= 65536, "n_samples must be in [1, 65536]");
    ADC_REQUIRE(a->n_grid > 0 && a->n_grid <= ADC_IDEAL_MAX_GRID && a->bid_grid_host != nullptr, "bid grid");
    ADC_REQUIRE(a->ideal_profit != nullptr, "ideal_profit is NULL");
    int rc = check_device();
    if (rc) return rc;
    if (a->device >= 0) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != a->device)
            return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: the buffers live on device %d but the "
                        "calling thread's current device is %d", a->device, cur);
    }
    const cudaError_t e = adc::launch_ideal_profit(*a, static_cast<cudaStream_t>(stream), &g_launches);
    if (e == cudaErrorInvalidValue)
        return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: a grid bid is NaN or not below 5.10");
    if (e != cudaSuccess) return fail(ADC_ERR_CUDA, "adcraft_b200: launch failed: %s", cudaGetErrorString(e));
    return ADC_OK;
}

int adc_sizeof_metrics_args(void) { return (int)sizeof(adc_metrics_args); }

int adc_episode_metrics(const adc_metrics_args *a, void *stream)
{
    if (a == nullptr) return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: args is NULL");
    ADC_REQUIRE(a->E > 0 && a->K > 0 && a->K <= ADC_METRICS_MAX_K, "E > 0 and 0 < K <= ADC_METRICS_MAX_K");
    ADC_REQUIRE(a->steps > 0, "steps must be positive");
    ADC_REQUIRE(a->episode_profit_cents && a->ideal && a->sums, "episode_profit_cents, ideal and sums are required");
    ADC_REQUIRE(a->ideal_env_stride == 0 || a->ideal_env_stride == a->K, "ideal_env_stride must be 0 or K");
    int rc = check_device();
    if (rc) return rc;
    if (a->device >= 0) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != a->device)
            return fail(ADC_ERR_INVALID, "adcraft_b200: invalid argument: the buffers live on device %d but the "
                        "calling thread's current device is %d", a->device, cur);
    }
    const cudaError_t e = adc::launch_episode_metrics(*a, static_cast<cudaStream_t>(stream), &g_launches);
    if (e != cudaSuccess) return fail(ADC_ERR_CUDA, "adcraft_b200: launch failed: %s", cudaGetErrorString(e));
    return ADC_OK;
}

int64_t adc_launch_count(int reset)
{
    const int64_t n = g_launches;
    if (reset) g_launches = 0;
    return n;
}

}  // extern "C"
