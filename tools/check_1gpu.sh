#!/bin/bash
# One-GPU check (run under gpurun): micro-benchmarks, GPU parity tests, smoke, both bench arms and the ncu launch list of
# one bench step.  Everything lands in gpurun_out/.   usage: bash tools/check_1gpu.sh <tag> [what...]
# what: micro pytest smoke bench ref launches (default: all)
tag=${1:-r02a}; shift
what=${*:-micro pytest smoke bench ref launches}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$tag.txt 2>&1
for w in $what; do case $w in
micro)
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lds_ffma_bench tools/micro/lds_ffma_bench.cu 2>/dev/null \
    && timeout 60 /tmp/lds_ffma_bench > $O/micro_$tag.log 2>&1; echo "== micro"; cat $O/micro_$tag.log;;
pytest)
  ( time timeout 700 python -m pytest tests -m gpu -x -q -n 4 --durations=12 ) > $O/pytest_gpu_$tag.log 2>&1
  echo "== pytest"; tail -25 $O/pytest_gpu_$tag.log;;
smoke)
  ( time timeout 150 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke_$tag.log 2>&1
  echo "== smoke"; tail -4 $O/smoke_$tag.log;;
bench)
  timeout 420 python bench.py > $O/bench_$tag.json 2> $O/bench_err_$tag.log
  echo "== bench"; tail -c 4000 $O/bench_$tag.json; tail -3 $O/bench_err_$tag.log;;
ref)
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$tag.json 2>> $O/bench_err_$tag.log
  echo "== reference"; tail -c 700 $O/bench_reference_$tag.json;;
launches)
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/launches_$tag.log 2>&1
  echo "== launches"; wc -l $O/launches_$tag.csv;;
esac; done
