"""The C-ABI library: loads without a GPU, exports every symbol include/adcraft_b200.h declares,
matches the ctypes layout, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "adcraft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from adcraft_b200 import _capi
    lib = _capi.load()
    declared = _declared_symbols()
    assert set(declared) == set(_capi.EXPORTED_SYMBOLS), declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_struct_layout_matches_compiled_library():
    from adcraft_b200 import _capi
    lib = _capi.load()
    header = open(os.path.join(ROOT, "include", "adcraft_b200.h")).read()
    declared = int(re.search(r"#define\s+ADC_ABI_VERSION\s+(\d+)", header).group(1))
    assert lib.adc_abi_version() == declared == _capi.ABI_VERSION
    assert lib.adc_sizeof_step_args() == C.sizeof(_capi.StepArgs)
    assert lib.adc_sizeof_tape() == C.sizeof(_capi.Tape)


def test_invalid_arguments_are_reported_not_crashed():
    from adcraft_b200 import _capi
    lib = _capi.load()
    a = _capi.StepArgs()
    rc = lib.adc_step_philox(C.byref(a), None)
    assert rc == -1 and b"invalid argument" in lib.adc_last_error()
    assert lib.adc_step_replay(C.byref(a), None, None) == -1
    assert lib.adc_reset_envs(0, None, None, None, None) == -1


def test_no_cpu_fallback_without_device():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from adcraft_b200 import _capi
    from adcraft_b200.vector_env import VectorBiddingSimulation
    assert _capi.load().adc_device_count() == 0
    with pytest.raises(_capi.AdcError, match="no CPU fallback"):
        VectorBiddingSimulation(4, num_keywords=3)
    # the C entry point itself also refuses: a fully valid argument block still fails with NO_DEVICE
    import numpy as np
    E, K = 2, 3
    f = lambda *s: np.zeros(s)
    keep = dict(kw=[f(K) for _ in range(8)], st=[f(E), f(E), np.zeros(E, np.int32)],
                bids=f(E, K), out=[np.zeros((E, K), np.int32) for _ in range(3)] + [f(E, K), f(E, K)],
                cents=[np.zeros((E, K), np.int64) for _ in range(2)],
                env=[f(E), f(E), np.zeros(E, np.int32), np.zeros(E, np.uint8), np.zeros(E, np.uint8)],
                sc=[np.zeros(E, np.int32), np.zeros(2, np.int32), np.zeros(E, np.int64),
                    np.zeros(E, np.int64), np.zeros(E, np.int32)])
    a = _capi.StepArgs()
    a.E, a.kw.K, a.kw.kind = E, K, 0
    for n, arr in zip(("vol_mean", "vol_std", "p1", "p2", "ctr", "cvr", "rev_mean", "rev_std"), keep["kw"]):
        setattr(a.kw, n, arr.ctypes.data)
    a.env.budget, a.env.cum_profit, a.env.day = [x.ctypes.data for x in keep["st"]]
    a.bids, a.bids_dtype = keep["bids"].ctypes.data, 1
    o = a.out
    o.impressions, o.clicks, o.conversions, o.cost, o.revenue = [x.ctypes.data for x in keep["out"]]
    o.float_dtype = 1
    o.cost_cents, o.revenue_cents = [x.ctypes.data for x in keep["cents"]]
    o.reward, o.obs_cum_profit, o.obs_days, o.terminated, o.truncated = [x.ctypes.data for x in keep["env"]]
    s = a.scratch
    s.serial_list, s.serial_count, s.env_profit, s.env_cost, s.env_done = [x.ctypes.data for x in keep["sc"]]
    rc = _capi.load().adc_step_philox(C.byref(a), None)
    assert rc == -2, _capi.load().adc_last_error()
    assert b"no CPU fallback" in _capi.load().adc_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from adcraft_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setenv("ADCRAFT_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_capi.AdcError, match="not found"):
        _capi.load()
