#!/bin/bash
# serial-walk iteration: parity suite, then the budget-bound bench lines (run under gpurun)
O=gpurun_out
timeout 800 python -m pytest tests -x -q -m gpu > $O/ser_tests.log 2>&1; echo "tests rc=$?" >> $O/ser_tests.log
tail -3 $O/ser_tests.log
A="--no-replay --no-cpu-baseline"
python bench.py $A --budget 1000 --steps 60 > $O/ser_b1000.json 2> $O/ser.err
python bench.py $A --budget 1000 --alias --steps 60 > $O/ser_b1000_alias.json 2>> $O/ser.err
python bench.py $A --keywords 1000 --envs 16384 --volume 16 --cvr 0.1 --budget 1000 --steps 30 > $O/ser_c3_b1000.json 2>> $O/ser.err
python - <<'PY'
import json
for f in ("ser_b1000", "ser_b1000_alias", "ser_c3_b1000"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"], 4), "%.3e" % d["value"])
    except Exception as ex: print(f, "ERR", ex)
PY
