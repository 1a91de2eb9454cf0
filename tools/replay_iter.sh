#!/bin/bash
# replay-kernel iteration: the replay parity tests, then the replay leg of the bench (run under gpurun)
O=gpurun_out
timeout 150 python -m pytest tests/test_gpu_replay_random.py tests/test_gpu_golden.py tests/test_gpu_outcomes.py -x -q -m gpu > $O/rp_tests.log 2>&1; echo "tests rc=$?" >> $O/rp_tests.log
tail -3 $O/rp_tests.log
timeout 90 python bench.py --replay-only --steps 60 --warmup 5 > $O/rp.json 2> $O/rp.err
python -c "
import json; d=json.loads(open('gpurun_out/rp.json').read().strip().splitlines()[-1])['replay']; print('replay ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'csr', round(d['csr_kernel']['ms_per_step'],4))"
