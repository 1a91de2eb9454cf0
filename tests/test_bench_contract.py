"""bench.py's output contract, checked on the CPU through the reference arm (the GPU arm prints the
same keys; it needs a device)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(env_extra=None, args=()):
    env = dict(os.environ, OMP_NUM_THREADS="2", **(env_extra or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", *args], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    return p.stdout.strip().splitlines()


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "keyword_auction_steps_per_sec" and d["unit"] == "keyword-auction-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("C2:") and d["config"]["envs_per_gpu"] == 4096
    cb = d["cpu_baseline"]
    # the reference's own Python when its staged copy (baseline/_ref) or /root/reference is present,
    # else the C port
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    if cb["kind"] == "reference":
        assert 50 < cb["units_per_s_per_core"] < 5000 and cb["c_port"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    assert _run({"RANK": "1", "WORLD_SIZE": "2"}, ("--gpus", "2")) == []
