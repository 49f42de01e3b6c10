#!/bin/bash
for d in 0 148 296 444 888 1776; do
  SMK_PF_DIST=$d python tools/bench_pass.py 512 512 1536 8 2>&1 | tail -2 | sed "s/^default/pf=$d/"
done
