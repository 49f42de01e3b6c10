#!/bin/bash
# A/B run of a kernel-tuning variant against the default build (run under gpurun): box parity tests + per-pass times.
#   here:  python -m saclaymocks_b200.build --variant <tag> <DEFINE=VALUE ...>     (libsmk_<tag>.so travels with the snapshot)
#   box:   bash tools/ab_check.sh saclaymocks_b200/libsmk_<tag>.so
# Used for SMK_Z_LINES (8- against 16-line z tiles), SMK_Z_TW (twiddle sources of the z pass) and SMK_VEL_F64.
for lib in "" "$@"; do
  export SMK_LIB_PATH=$lib; [ -z "$lib" ] && unset SMK_LIB_PATH
  echo "== lib ${lib:-default}"
  timeout 200 python -m pytest tests/test_gpu_boxes.py tests/test_gpu_sizes.py -m gpu -x -q -n 4 2>&1 | tail -1
  timeout 150 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step %.2f boxes %.2f gather %.2f' % (d['ms_per_step'], d['t_boxes_ms'], d['t_gather_ms']), {k: round(v, 3) for k, v in d['roofline']['passes_ms'].items()})"
done
