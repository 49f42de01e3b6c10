#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step, then --set full captures of the kernels of the step,
# exported to CSV on the GPU box (the .ncu-rep files together exceed what gpurun copies back).
# usage: bash tools/profile.sh <tag>     (run under gpurun, one GPU)
tag=${1:-r01b}
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu"
T=/tmp/smkprof; mkdir -p $T
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"skewers_|draw_qso|smallscale|fgpa" -c 4 -f -o $T/prof_spec_$tag $B > gpurun_out/prof_spec_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"c2c_strided|c2r_z|r2c_z" -c 18 -f -o $T/prof_fft_$tag $B > gpurun_out/prof_fft_$tag.log 2>&1
for k in spec fft; do
  ncu -i $T/prof_${k}_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${k}_$tag.csv 2>/dev/null
done
ncu -i $T/prof_spec_$tag.ncu-rep --page source --csv -k regex:skewers_tma > gpurun_out/ncu_source_skewers_$tag.csv 2>/dev/null
cp $T/prof_spec_$tag.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out/
