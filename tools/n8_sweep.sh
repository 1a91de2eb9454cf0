#!/bin/bash
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T tools/host_ceiling.py > $O/r2_host_ceiling_n8.json 2> $O/r2_n8.err
$T bench.py --gpus 8 --steps 240 --warmup 10 > $O/r2_bench_c2_n8.json 2>> $O/r2_n8.err
$T bench.py --gpus 8 --config c5 --steps 3 --warmup 1 > $O/r2_bench_c5_n8.json 2>> $O/r2_n8.err
python bench.py --config c5 --steps 3 --warmup 1 > $O/r2_bench_c5_n1.json 2>> $O/r2_n8.err
tail -5 $O/r2_n8.err
cat $O/r2_host_ceiling_n8.json
