#!/bin/bash
# z pass with the two-stage plan 768 = 32.24 and permuted read-back (libsmk_z2.so = -DSMK_PLAN_768=2) against the
# default 16.16.3 with re-sort: per-pass times, parity on the variant, one bench line.  One GPU.
O=gpurun_out; mkdir -p $O
P=$PWD/saclaymocks_b200
for lib in libsmk.so libsmk_z2.so; do
  echo "== $lib"; SMK_LIB_PATH=$P/$lib timeout 100 python tools/bench_pass.py 512 512 1536 8 2>&1 | tail -2
done | tee $O/z_sweep_passes.log
echo "== parity on the variant"
SMK_LIB_PATH=$P/libsmk_z2.so timeout 200 python -m pytest tests/test_gpu_boxes.py tests/test_gpu_sizes.py tests/test_gpu_statistics.py tests/test_gpu_spectra.py -m gpu -x -q -n 4 2>&1 | tail -4 | tee $O/z_sweep_pytest.log
SMK_LIB_PATH=$P/libsmk_z2.so timeout 100 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 > $O/bench_z2.json
python -c "
import json; d=json.load(open('$O/bench_z2.json')); print('z2: step', d['ms_per_step'], 'boxes', d['t_boxes_ms'], 'skewers', d['t_skewers_ms'], d['roofline']['passes_ms'])"
