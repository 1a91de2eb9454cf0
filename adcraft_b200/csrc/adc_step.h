// Internal launch interface between the kernels (adc_step.cu) and the C ABI (adc_capi.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/adcraft_b200.h"

namespace adc {

// Enqueue one env step on `s`; `tape` == nullptr selects the free-running (Philox) sources.
cudaError_t launch_step(const adc_step_args &a, const adc_tape *tape, cudaStream_t s, int64_t *launches);
cudaError_t launch_reset_envs(int32_t E, const uint8_t *mask, double *cum_profit, int32_t *day,
                              cudaStream_t s, int64_t *launches);

int64_t serial_slab_bytes(int32_t K);
int64_t host_row_bytes(int32_t K, int32_t float_dtype);
cudaError_t launch_pack_rows(const adc_step_args &a, void *rows_dev, cudaStream_t s, int64_t *launches);
cudaError_t launch_ideal_profit(const adc_ideal_args &a, cudaStream_t s, int64_t *launches);
cudaError_t launch_episode_metrics(const adc_metrics_args &a, cudaStream_t s, int64_t *launches);

}  // namespace adc
