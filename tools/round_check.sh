#!/bin/bash
# End-of-round check on one B200 (run under gpurun): GPU parity tests, smoke, the bench line of both arms, the ncu
# launch list of one bench step, and the skewer prefetch-distance sweep.  Everything lands in gpurun_out/.
# usage: bash tools/round_check.sh <tag>
tag=${1:-r01c}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu_$tag.txt 2>&1
( time timeout 560 python -m pytest tests -m gpu -x -q -n 4 --durations=12 ) > $O/pytest_gpu_$tag.log 2>&1
echo "== pytest"; tail -22 $O/pytest_gpu_$tag.log
( time timeout 150 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke_$tag.log 2>&1
echo "== smoke"; tail -4 $O/smoke_$tag.log
timeout 420 python bench.py > $O/bench_$tag.json 2> $O/bench_err_$tag.log
echo "== bench"; tail -c 3000 $O/bench_$tag.json; tail -3 $O/bench_err_$tag.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/launches_$tag.log 2>&1
echo "== launches"; wc -l $O/launches_$tag.csv
timeout 240 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$tag.json 2>> $O/bench_err_$tag.log
echo "== reference"; tail -c 600 $O/bench_reference_$tag.json
echo "== pfd sweep"
if [ -n "$PFDS" ]; then timeout 200 bash tools/skew_pfd_sweep.sh 2>&1 | tee $O/pfd_sweep_$tag.log; fi
if [ "${FULL_NCU:-0}" = "1" ]; then echo "== ncu --set full"; bash tools/profile.sh $tag > $O/profile_$tag.log 2>&1; tail -12 $O/profile_$tag.log; fi
