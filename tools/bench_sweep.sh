#!/bin/bash
# measurement sweep used for profiles/r02_*.json (run under gpurun)
set -x
O=gpurun_out
python bench.py --steps 240 --warmup 10 > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err
python bench.py --budget 1000 --no-replay --no-cpu-baseline --steps 60 > $O/r2_bench_budget1000.json 2>> $O/r2_bench_c2.err
python bench.py --budget 1000 --alias --no-replay --no-cpu-baseline --steps 60 > $O/r2_bench_budget1000_alias.json 2>> $O/r2_bench_c2.err
python bench.py --keywords 1000 --envs 16384 --volume 16 --cvr 0.1 --no-replay --no-cpu-baseline --steps 60 > $O/r2_bench_c3_sparse.json 2>> $O/r2_bench_c2.err
python bench.py --keywords 1000 --envs 16384 --volume 64 --cvr 0.1 --drift --no-replay --no-cpu-baseline --steps 60 > $O/r2_bench_c3_nonstat.json 2>> $O/r2_bench_c2.err
python bench.py --keywords 1000 --envs 16384 --volume 16 --cvr 0.1 --budget 1000 --no-replay --no-cpu-baseline --steps 30 > $O/r2_bench_c3_sparse_budget1000.json 2>> $O/r2_bench_c2.err
python bench.py --config c4 --steps 20 > $O/r2_bench_c4.json 2>> $O/r2_bench_c2.err
tail -5 $O/r2_bench_c2.err
