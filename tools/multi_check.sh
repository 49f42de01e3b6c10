#!/bin/bash
# 2-GPU check (gpurun --gpus 2): sharded parity tests (both exchanges, persistent fused x pass), the bench line at N=2
# with its end-to-end arm, and the persistent x pass (SMK_X_SMS) against the default.
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 240 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > $O/pytest_multi.log 2>&1; tail -6 $O/pytest_multi.log
timeout 200 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; tail -c 2500 $O/bench_2gpu.json; tail -3 $O/bench_2gpu.err
for x in 32 64; do
  SMK_X_SMS=$x timeout 120 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 > $O/bench_2gpu_x$x.json
  python -c "
import json; d=json.load(open('$O/bench_2gpu_x$x.json')); print('SMK_X_SMS=$x step', d['ms_per_step'], 'boxes', d['t_boxes_ms'], 'skewers', d['t_skewers_ms'])"
done
