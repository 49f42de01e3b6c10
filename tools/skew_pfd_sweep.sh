#!/bin/bash
# skewer gather: distance (rows) of the L1 row prefetch
for d in ${PFDS:-1 2 3 4}; do
  SMK_SKEW_PFD=$d python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pfd $d skewers', d['t_skewers_ms'], 'boxes', d['t_boxes_ms'], 'qso kernel', d['t_draw_qso_kernel_ms'])"
done
