#!/bin/bash
# Two-stage radix-32 plans for the strided passes (libsmk_p2.so = -DSMK_PLAN_512=2 -DSMK_PLAN_1024=2, the default since round 1; the old plans are =3) against the
# default library: per-pass times, parity tests on the variant, one bench line.  Run under gpurun (one GPU).
O=gpurun_out; mkdir -p $O
P=$PWD/saclaymocks_b200
for lib in libsmk.so libsmk_p2.so; do
  for n in 512 1024; do
    echo "== $lib $n"; SMK_LIB_PATH=$P/$lib timeout 120 python tools/bench_pass.py $n $n 1536 6 2>&1 | tail -2
  done
done | tee $O/plan_sweep_passes.log
echo "== parity on the variant"
SMK_LIB_PATH=$P/libsmk_p2.so timeout 300 python -m pytest tests/test_gpu_boxes.py tests/test_gpu_sizes.py tests/test_gpu_statistics.py -m gpu -x -q -n 4 2>&1 | tail -5 | tee $O/plan_sweep_pytest.log
echo "== bench on the variant"
SMK_LIB_PATH=$P/libsmk_p2.so timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 > $O/bench_p2.json
python -c "
import json; d=json.load(open('$O/bench_p2.json')); print('p2: step', d['ms_per_step'], 'boxes', d['t_boxes_ms'], 'skewers', d['t_skewers_ms'], d['roofline']['passes_ms'])"
echo "== default lib: fixed test"
timeout 200 python -m pytest tests/test_gpu_sizes.py -m gpu -x -q -k run_chunk 2>&1 | tail -3
