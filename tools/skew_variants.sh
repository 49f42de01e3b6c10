#!/bin/bash
# staged skewer gather: pixels per thread x fields per stage (SMK_SKEW_VARIANT = 10 P + NFG), one bench step each
for v in ${*:-42 41 61 81}; do
  SMK_SKEW_VARIANT=$v timeout 150 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v gather %.2f ms  skewers %.2f  boxes %.2f  c2r_z %.3f  staging %s' % (d['t_gather_ms'], d['t_skewers_ms'], d['t_boxes_ms'], d['roofline']['passes_ms']['c2r_z'], d['gather_staging_rank0']))"
done
