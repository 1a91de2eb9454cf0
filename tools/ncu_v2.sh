#!/bin/bash
set -x
O=gpurun_out
A="--steps 3 --warmup 3 --no-replay --no-cpu-baseline"
python bench.py $A > $O/ncu_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:adc_flat2 -s 4 -c 1 -o $O/r2_flat2_c2 python bench.py $A > $O/ncu_c2.log 2>&1
python bench.py $A --budget 1000 > $O/ncu_plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:adc_serial_warp -s 4 -c 1 -o $O/r2_serial_b1000 python bench.py $A --budget 1000 > $O/ncu_b.log 2>&1
ls -la $O/*.ncu-rep
