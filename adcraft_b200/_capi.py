"""ctypes binding of include/adcraft_b200.h (the C ABI of the CUDA library).

The product path has NO fallback: if the shared library is missing, or there is no CUDA
device, loading / stepping raises.  Layout is verified against the compiled struct sizes.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

SUBSTEPS = 24
IMPLICIT, EXPLICIT, IMPLICIT_MULTI = 0, 1, 2
F32, F64 = 0, 1
ABI_VERSION = 6


class AdcError(RuntimeError):
    pass


class Keywords(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("K", C.c_int32), ("env_stride", C.c_int64),
        ("vol_mean", C.c_void_p), ("vol_std", C.c_void_p), ("p1", C.c_void_p), ("p2", C.c_void_p),
        ("ctr", C.c_void_p), ("cvr", C.c_void_p), ("rev_mean", C.c_void_p), ("rev_std", C.c_void_p),
        ("max_bidders", C.c_void_p), ("participation", C.c_void_p),
        ("impression_thresh", C.c_double),
    ]


class EnvState(C.Structure):
    _fields_ = [
        ("budget", C.c_void_p), ("cum_profit", C.c_void_p), ("day", C.c_void_p),
        ("max_days", C.c_int32), ("loss_threshold", C.c_double),
    ]


class Drift(C.Structure):
    _fields_ = [("mask", C.c_void_p), ("num_updates", C.c_int32), ("mag", C.c_double * 3)]


class StepOut(C.Structure):
    _fields_ = [
        ("impressions", C.c_void_p), ("clicks", C.c_void_p), ("conversions", C.c_void_p),
        ("cost", C.c_void_p), ("revenue", C.c_void_p), ("float_dtype", C.c_int32),
        ("cost_cents", C.c_void_p), ("revenue_cents", C.c_void_p), ("reward", C.c_void_p),
        ("obs_cum_profit", C.c_void_p), ("obs_days", C.c_void_p), ("terminated", C.c_void_p),
        ("truncated", C.c_void_p), ("remaining_budget", C.c_void_p),
        ("episode_profit_cents", C.c_void_p), ("episode_reward", C.c_void_p), ("episode_count", C.c_void_p),
        ("rows", C.c_void_p), ("flat_obs", C.c_void_p), ("unit_records", C.c_void_p),
    ]


class Scratch(C.Structure):
    _fields_ = [
        ("serial_list", C.c_void_p), ("serial_count", C.c_void_p), ("env_profit", C.c_void_p),
        ("env_cost", C.c_void_p), ("env_done", C.c_void_p), ("unit_cost_f64", C.c_void_p),
        ("work_counter", C.c_void_p), ("acc_impressions", C.c_void_p), ("acc_clicks", C.c_void_p), ("acc_conversions", C.c_void_p),
        ("serial_ws", C.c_void_p), ("serial_ws_bytes", C.c_int64),
        ("serial_hint", C.c_void_p), ("outbid_mask", C.c_void_p),
    ]


class Detail(C.Structure):
    _fields_ = [("cap", C.c_int32), ("costs", C.c_void_p), ("rev_per_cost", C.c_void_p),
                ("n_recorded", C.c_void_p), ("volume_seen", C.c_void_p),
                ("lane_clicks", C.c_void_p), ("lane_convs", C.c_void_p)]


class StepArgs(C.Structure):
    _fields_ = [
        ("E", C.c_int32), ("env_base", C.c_uint32), ("step", C.c_uint32), ("parity", C.c_uint32),
        ("device", C.c_int32), ("seed", C.c_uint64),
        ("n_lanes", C.c_int32), ("budget_alias", C.c_int32), ("autoreset", C.c_int32),
        ("force_serial", C.c_int32),
        ("kw", Keywords), ("env", EnvState), ("drift", Drift),
        ("bids", C.c_void_p), ("bids_dtype", C.c_int32), ("f32_ties", C.c_int32),
        ("budget_in", C.c_void_p),
        ("env_group", C.c_int32), ("spread_outcomes", C.c_int32), ("floor_cents", C.c_void_p),
        ("out", StepOut), ("scratch", Scratch), ("detail", Detail),
    ]


class HostChunk(C.Structure):
    _fields_ = [("args", StepArgs), ("bids_host", C.c_void_p), ("rows_dev", C.c_void_p),
                ("rows_host", C.c_void_p), ("stream", C.c_void_p)]


class IdealArgs(C.Structure):
    _fields_ = [
        ("E", C.c_int32), ("env_base", C.c_uint32), ("step", C.c_uint32), ("device", C.c_int32),
        ("seed", C.c_uint64), ("kw", Keywords), ("n_samples", C.c_int32), ("n_grid", C.c_int32),
        ("bid_grid_host", C.c_void_p), ("samples_cents", C.c_void_p), ("ideal_profit", C.c_void_p),
        ("positive_frac", C.c_void_p), ("best_bid_index", C.c_void_p), ("impression_rate", C.c_void_p),
        ("expected_cpc", C.c_void_p),
    ]


class MetricsArgs(C.Structure):
    _fields_ = [
        ("E", C.c_int32), ("K", C.c_int32), ("steps", C.c_int32), ("device", C.c_int32),
        ("episode_profit_cents", C.c_void_p), ("ideal", C.c_void_p), ("ideal_env_stride", C.c_int64),
        ("sums", C.c_void_p), ("akncp", C.c_void_p), ("ncp", C.c_void_p), ("zero", C.c_int32), ("pad_", C.c_int32),
    ]


METRICS_MAX_K = 2048


class Tape(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "volume", "comp_off", "comp_cents", "comp_f64", "click_off", "u_click", "conv_off", "u_conv",
        "rev_off", "rev_cents", "impr", "cost_off", "cost", "drift", "packed", "packed_off")]


_lib = None


def library_path() -> str:
    return os.environ.get("ADCRAFT_B200_LIB", _build.LIB_PATH)


def load() -> C.CDLL:
    """Load the CUDA library; raise if it is missing or its ABI does not match."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise AdcError(
            f"adcraft_b200: CUDA library not found at {path}. Build it with "
            "`python -m adcraft_b200.build` (needs nvcc); there is no CPU fallback.")
    lib = C.CDLL(path)
    lib.adc_last_error.restype = C.c_char_p
    lib.adc_abi_version.restype = C.c_int
    lib.adc_device_count.restype = C.c_int
    lib.adc_sizeof_step_args.restype = C.c_int
    lib.adc_sizeof_tape.restype = C.c_int
    lib.adc_step_philox.restype = C.c_int
    lib.adc_step_philox.argtypes = [C.POINTER(StepArgs), C.c_void_p]
    lib.adc_step_replay.restype = C.c_int
    lib.adc_step_replay.argtypes = [C.POINTER(StepArgs), C.POINTER(Tape), C.c_void_p]
    lib.adc_reset_envs.restype = C.c_int
    lib.adc_reset_envs.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.adc_ideal_profit.restype = C.c_int
    lib.adc_ideal_profit.argtypes = [C.POINTER(IdealArgs), C.c_void_p]
    lib.adc_sizeof_ideal_args.restype = C.c_int
    lib.adc_episode_metrics.restype = C.c_int
    lib.adc_episode_metrics.argtypes = [C.POINTER(MetricsArgs), C.c_void_p]
    lib.adc_sizeof_metrics_args.restype = C.c_int
    lib.adc_host_row_bytes.restype = C.c_int64
    lib.adc_host_row_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.adc_step_host.restype = C.c_int
    lib.adc_step_host.argtypes = [C.POINTER(HostChunk), C.c_int32]
    lib.adc_sizeof_host_chunk.restype = C.c_int
    lib.adc_serial_slab_bytes.restype = C.c_int64
    lib.adc_serial_slab_bytes.argtypes = [C.c_int32]
    lib.adc_launch_count.restype = C.c_int64
    lib.adc_launch_count.argtypes = [C.c_int]
    if lib.adc_abi_version() != ABI_VERSION:
        raise AdcError(f"adcraft_b200: ABI version {lib.adc_abi_version()} != {ABI_VERSION}")
    if lib.adc_sizeof_step_args() != C.sizeof(StepArgs) or lib.adc_sizeof_tape() != C.sizeof(Tape):
        raise AdcError(
            "adcraft_b200: struct layout mismatch between _capi.py and the compiled library "
            f"({lib.adc_sizeof_step_args()} vs {C.sizeof(StepArgs)}, "
            f"{lib.adc_sizeof_tape()} vs {C.sizeof(Tape)}); rebuild")
    if lib.adc_sizeof_host_chunk() != C.sizeof(HostChunk):
        raise AdcError("adcraft_b200: adc_host_chunk layout mismatch; rebuild")
    if lib.adc_sizeof_ideal_args() != C.sizeof(IdealArgs):
        raise AdcError("adcraft_b200: adc_ideal_args layout mismatch; rebuild")
    if lib.adc_sizeof_metrics_args() != C.sizeof(MetricsArgs):
        raise AdcError("adcraft_b200: adc_metrics_args layout mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().adc_last_error()
        raise AdcError(f"{msg.decode() if msg else 'adcraft_b200 error'} (status {rc})")


EXPORTED_SYMBOLS = (
    "adc_last_error", "adc_abi_version", "adc_device_count", "adc_sizeof_step_args",
    "adc_sizeof_tape", "adc_step_philox", "adc_step_replay", "adc_reset_envs", "adc_launch_count",
    "adc_ideal_profit", "adc_sizeof_ideal_args", "adc_serial_slab_bytes",
    "adc_host_row_bytes", "adc_step_host", "adc_sizeof_host_chunk",
    "adc_episode_metrics", "adc_sizeof_metrics_args",
)
