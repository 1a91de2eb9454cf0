#!/bin/bash
# round-2 measurement + profile capture (run under gpurun on one B200)
O=gpurun_out
A="--steps 3 --warmup 3 --no-replay --no-cpu-baseline"
python bench.py $A > $O/ncu_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:adc_flat2 -s 4 -c 1 -f -o $O/r2_flat2_c2 python bench.py $A > $O/ncu_c2.log 2>&1
python bench.py $A --budget 1000 > $O/ncu_plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:adc_serial_warp -s 4 -c 1 -f -o $O/r2_serial_b1000 python bench.py $A --budget 1000 > $O/ncu_b.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/ncu_summary.py $O/r2_flat2_c2.ncu-rep $O/r02_flat2_c2_ncu --envs 4096 --keywords 100
python tools/ncu_summary.py $O/r2_serial_b1000.ncu-rep $O/r02_serial_warp_b1000_ncu --envs 4096 --keywords 100 --note "C2 with budget 1000: every env binds"
cp $O/r02_flat2_c2_ncu.json profiles/r02_flat2_c2_ncu.json
ncu -i $O/r2_flat2_c2.ncu-rep --page source --print-source cuda,sass --csv > $O/r2_flat2_src.csv 2>/dev/null
ncu -i $O/r2_serial_b1000.ncu-rep --page source --print-source cuda,sass --csv > $O/r2_serial_src.csv 2>/dev/null

bash tools/bench_sweep.sh > $O/sweep.log 2>&1
python bench.py --explicit --steps 20 --no-replay --no-cpu-baseline > $O/r2_bench_explicit.json 2>> $O/sweep.log
ls $O | tail -30
# replay (packed tape) kernel: one full capture + summary
python bench.py --replay-only --steps 3 --warmup 3 > $O/ncu_plain_replay.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:adc_replay_packed -s 4 -c 1 -f -o $O/r2_replay_packed python bench.py --replay-only --steps 3 --warmup 3 > $O/ncu_replay.log 2>&1
python tools/ncu_summary.py $O/r2_replay_packed.ncu-rep $O/r02_replay_packed_ncu --envs 4096 --keywords 100 --note "tape-driven step on the packed tape (bench.py --replay-only)"

rm -f $O/*.ncu-rep
