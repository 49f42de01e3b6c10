#!/bin/bash
# skewer gather: FFMA2 variant x prefetch mode (0 none, 1 L2 next x slab, 2 L1 next row, 3 both)
for pf in ${PFS:-0 1 2 3}; do
  SMK_SKEW_VARIANT=42 SMK_SKEW_PF=$pf python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant 42 pf $pf skewers', d['t_skewers_ms'], 'boxes', d['t_boxes_ms'], d['roofline']['passes_ms'])"
done
