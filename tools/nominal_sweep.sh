#!/bin/bash
# 8 GPUs (gpurun --gpus 8): the nominal 2560 x 2560 x 1536 chunk with the fused-exchange x pass persistent on
# SMK_X_SMS CTAs (0 = one CTA per tile, the default), then the weak-scaling box of bench.py at N=8.
# What to look for: boxes ms of the nominal chunk against the serial HBM+NVLink roofline of ~136 ms.
O=gpurun_out; mkdir -p $O
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for x in ${XSMS:-0 48 64 96}; do
  SMK_X_SMS=$x timeout 300 $TR bench.py --gpus $N --box 2560 --steps 3 --warmup 2 --no-cpu 2>/dev/null | grep '^{"metric' > $O/bench_nominal_x$x.json
  python -c "
import json; d=json.load(open('$O/bench_nominal_x$x.json')); print('nominal SMK_X_SMS=$x step', d['ms_per_step'], 'boxes', d['t_boxes_ms'], 'skewers', d['t_skewers_ms'])"
done | tee $O/nominal_sweep.log
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 | grep '^{"metric' > $O/bench_${N}gpu.json; tail -c 400 $O/bench_${N}gpu.json
