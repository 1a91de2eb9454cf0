#!/bin/bash
# ncu capture of one launch of the exact serial walk: C2 with budget 1000 (default) or `c3` (sparse 1000 x 16384)
O=gpurun_out
if [ "$1" = "c3" ]; then
  A="--steps 3 --warmup 3 --no-replay --no-cpu-baseline --budget 1000 --keywords 1000 --envs 16384 --volume 16 --cvr 0.1"; TAG=c3_b1000; EK="--envs 16384 --keywords 1000"
else
  A="--steps 3 --warmup 3 --no-replay --no-cpu-baseline --budget 1000"; TAG=b1000; EK="--envs 4096 --keywords 100"
fi
python bench.py $A > $O/ncu_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:adc_serial_warp -s 4 -c 1 -f -o $O/r2_serial_$TAG python bench.py $A > $O/ncu_$TAG.log 2>&1
python tools/ncu_summary.py $O/r2_serial_$TAG.ncu-rep $O/r02_serial_warp_${TAG}_ncu $EK --note "budget 1000: every env binds"
ncu -i $O/r2_serial_$TAG.ncu-rep --page source --print-source cuda,sass --csv > $O/r2_serial_${TAG}_src.csv 2>/dev/null
rm -f $O/r2_serial_$TAG.ncu-rep
