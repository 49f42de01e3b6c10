#!/bin/bash
for v in 4 44 3 2; do
  SMK_SKEW_VARIANT=$v python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', d['t_skewers_ms'])"
done
