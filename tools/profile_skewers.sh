#!/bin/bash
# ncu --set full capture of the skewer gather kernels of one bench step (one GPU, under gpurun); raw + source pages
# exported to CSV here because the .ncu-rep may exceed what gpurun copies back.   usage: bash tools/profile_skewers.sh <tag>
tag=${1:-r02b}
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
T=/tmp/smkprof; mkdir -p $T gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"skewers_tma|skewers_multi" -c 4 -f -o $T/prof_skew_$tag $B > gpurun_out/prof_skew_$tag.log 2>&1
ncu -i $T/prof_skew_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_skew_$tag.csv 2>/dev/null
ncu -i $T/prof_skew_$tag.ncu-rep --page source --csv -k regex:skewers_tma > gpurun_out/ncu_source_skew_$tag.csv 2>/dev/null
ncu -i $T/prof_skew_$tag.ncu-rep --page details -k regex:skewers_tma 2>/dev/null | head -150 > gpurun_out/ncu_details_skew_$tag.txt
ls -la $T gpurun_out | tail -8
