#!/bin/bash
# y<->z chaining through L2: planes per group x streams (x discard / persist) on one 512 x 512 x 1536 box
python tools/bench_pass.py 512 512 1536 8 2>&1 | tail -2
for g in 2 4 8 16 32; do for s in 1 2 3; do
  SMK_YZ_GROUP=$g SMK_YZ_STREAMS=$s python tools/bench_pass.py 512 512 1536 8 2>&1 | tail -2
done; done
echo "--- no discard"
for g in 4 8; do SMK_YZ_DISCARD=0 SMK_YZ_GROUP=$g SMK_YZ_STREAMS=2 python tools/bench_pass.py 512 512 1536 8 2>&1 | tail -2; done
echo "--- persist window"
for g in 4 8 16; do SMK_YZ_PERSIST=1 SMK_YZ_GROUP=$g SMK_YZ_STREAMS=2 python tools/bench_pass.py 512 512 1536 8 2>&1 | tail -2; done
