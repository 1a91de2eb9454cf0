#!/bin/bash
O=gpurun_out
A="--steps 3 --warmup 3 --no-replay --no-cpu-baseline --budget 1000"
python bench.py $A > $O/ncu_plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:adc_serial_warp -s 4 -c 1 -f -o $O/r2_serial_b1000 python bench.py $A > $O/ncu_b.log 2>&1
python tools/ncu_summary.py $O/r2_serial_b1000.ncu-rep $O/r02_serial_warp_b1000_ncu --envs 4096 --keywords 100 --note "C2 with budget 1000: every env binds"
ncu -i $O/r2_serial_b1000.ncu-rep --page source --print-source cuda,sass --csv > $O/r2_serial_src.csv 2>/dev/null
rm -f $O/r2_serial_b1000.ncu-rep
