#!/bin/bash
# Times the C2 step at several per-day budgets (binding budgets route envs through the exact serial kernel).
# usage: tools/budget_sweep.sh [extra bench.py flags]
for b in 1000 1800 2200 100000; do
  timeout 120 python bench.py --budget $b --steps 60 --warmup 5 --no-replay --no-cpu-baseline "$@" 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('budget', $b, 'ms_per_step', round(d['ms_per_step'],4))"
done
