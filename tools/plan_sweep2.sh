#!/bin/bash
# strided-pass variants at the multi-GPU transform lengths, measured on one GPU with thin boxes:
#   p2 = two-stage 512/1024 plans, p3 = p2 + 8-column tiles at 1024, p4 / p5 = p2 + 2560 = 32.16.5 / 16.32.5
O=gpurun_out; mkdir -p $O
P=$PWD/saclaymocks_b200
{
for lib in libsmk_p2.so libsmk_p3.so; do
  echo "== $lib 1024x1024"; SMK_LIB_PATH=$P/$lib timeout 120 python tools/bench_pass.py 1024 1024 1536 5 2>&1 | tail -2
done
for lib in libsmk_p2.so libsmk_p4.so libsmk_p5.so; do
  echo "== $lib 2560x256 (x pass 2560)"; SMK_LIB_PATH=$P/$lib timeout 120 python tools/bench_pass.py 2560 256 1536 5 2>&1 | tail -2
  echo "== $lib 256x2560 (y pass 2560)"; SMK_LIB_PATH=$P/$lib timeout 120 python tools/bench_pass.py 256 2560 1536 5 2>&1 | tail -2
done
} | tee $O/plan_sweep2_passes.log
for lib in libsmk_p3.so libsmk_p4.so libsmk_p5.so; do
  echo "== parity $lib"
  SMK_LIB_PATH=$P/$lib timeout 200 python -m pytest tests/test_gpu_sizes.py -m gpu -x -q -n 4 -k long_axis 2>&1 | tail -2
done | tee $O/plan_sweep2_pytest.log
