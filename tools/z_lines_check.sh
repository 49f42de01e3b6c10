#!/bin/bash
# z passes with 8-line tiles (4 CTAs of 128 threads per SM) against the default 16-line tiles: parity + per-pass times
for lib in "" saclaymocks_b200/libsmk_z8.so; do
  export SMK_LIB_PATH=$lib; [ -z "$lib" ] && unset SMK_LIB_PATH
  echo "== lib ${lib:-default}"
  timeout 200 python -m pytest tests/test_gpu_boxes.py -m gpu -x -q -n 4 2>&1 | tail -1
  timeout 150 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step %.2f boxes %.2f gather %.2f' % (d['ms_per_step'], d['t_boxes_ms'], d['t_gather_ms']), {k: round(v, 3) for k, v in d['roofline']['passes_ms'].items()})"
done
