#!/bin/bash
# fused-multiply x pass: first-stage batch / resident CTAs (libsmk variants built with saclaymocks_b200.build --variant)
for lib in libsmk.so libsmk_b1.so libsmk_oldx.so; do
  echo "== $lib"; SMK_LIB_PATH=$PWD/saclaymocks_b200/$lib python tools/bench_pass.py 512 512 1536 8 2>&1 | tail -2
done
for v in 4 42 442; do
  SMK_SKEW_VARIANT=$v python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skew variant $v', d['t_skewers_ms'], 'boxes', d['t_boxes_ms'])"
done
