#!/bin/bash
# One ncu --set full capture of the kernels matching <regex> in a short bench run; details / raw / source pages are
# exported on the box into gpurun_out/ (run under gpurun, one GPU).
# usage: bash tools/profile_kernel.sh <regex> <tag> [launches=2]
re=$1; tag=$2; n=${3:-2}
T=/tmp/smkprof; mkdir -p $T gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$re" -c $n -f -o $T/k_$tag \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-selfcheck > gpurun_out/prof_$tag.log 2>&1
ncu -i $T/k_$tag.ncu-rep --page details > gpurun_out/ncu_details_$tag.txt 2>/dev/null
ncu -i $T/k_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$tag.csv 2>/dev/null
ncu -i $T/k_$tag.ncu-rep --page source --csv > gpurun_out/ncu_source_$tag.csv 2>/dev/null
ls -la gpurun_out/ | grep $tag
