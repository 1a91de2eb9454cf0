#!/bin/bash
# serial-walk iteration: parity suite, then the budget-bound bench lines
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/ser_tests.log 2>&1; echo "tests rc=$?" >> $O/ser_tests.log
tail -5 $O/ser_tests.log
B="--no-replay --no-cpu-baseline --steps 60"
timeout 300 python bench.py --budget 1000 $B > $O/ser_b1000.json 2> $O/ser.err
timeout 300 python bench.py --budget 1000 --alias $B > $O/ser_b1000_alias.json 2>> $O/ser.err
timeout 300 python bench.py --keywords 1000 --envs 16384 --volume 16 --cvr 0.1 --budget 1000 --no-replay --no-cpu-baseline --steps 30 > $O/ser_c3_b1000.json 2>> $O/ser.err
python - <<'PY'
import json
for n in ("ser_b1000","ser_b1000_alias","ser_c3_b1000"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1]); print(n, d["ms_per_step"], d["value"], d.get("config",{}).get("serial_envs_per_step"))
    except Exception as ex: print(n, "ERR", ex)
PY
tail -3 $O/ser.err
