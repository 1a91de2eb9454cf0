#!/bin/bash
# hot-kernel iteration: GPU suite, then the free-running bench lines (run under gpurun)
O=gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > $O/hot_tests.log 2>&1; echo "tests rc=$?" >> $O/hot_tests.log
tail -2 $O/hot_tests.log
A="--no-replay --no-cpu-baseline"
timeout 120 python bench.py $A --steps 240 --warmup 10 > $O/hot_c2.json 2> $O/hot.err
timeout 120 python bench.py $A --keywords 1000 --envs 16384 --volume 16 --cvr 0.1 --steps 60 > $O/hot_c3.json 2>> $O/hot.err
timeout 120 python bench.py $A --keywords 1000 --envs 16384 --volume 64 --cvr 0.1 --drift --steps 60 > $O/hot_c3ns.json 2>> $O/hot.err
python - <<'PY'
import json
for f in ("hot_c2", "hot_c3", "hot_c3ns"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"], 4), "%.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"])
    except Exception as ex: print(f, "ERR", ex)
PY
