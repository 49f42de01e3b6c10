#!/bin/bash
# Multi-GPU check (gpurun --gpus N): sharded parity tests for the rank counts that fit, then the bench line at N GPUs
# (BASELINE configuration of that rank count, parity self-check, end-to-end arm).   usage: bash tools/multi_check.sh <N> <tag> [extra bench args]
N=${1:-2}; tag=${2:-r02m}; shift; shift
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
sel="w2-"; [ "$N" -ge 4 ] && sel="w2- or w4-"; [ "$N" -ge 8 ] && sel="w4- or w8-"
( time timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "$sel" ) > $O/pytest_multi_${N}gpu_$tag.log 2>&1
grep -E "passed|failed|skipped|^FAILED|^E  " $O/pytest_multi_${N}gpu_$tag.log | cut -c1-250 | head -20
timeout 600 $TR bench.py --gpus $N --steps 4 --warmup 3 "$@" > $O/bench_${N}gpu_$tag.json 2> $O/bench_${N}gpu_$tag.err
tail -4 $O/bench_${N}gpu_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_${N}gpu_$tag.json").read().strip().splitlines()[-1])
    print("N=$N box", d["config"]["box"], "step %.2f ms  boxes %.2f  skewers %.2f  gather %.2f" % (d["ms_per_step"], d["t_boxes_ms"], d["t_skewers_ms"], d["t_gather_ms"]))
    print("value %.3e  e2e %s" % (d["value"], d["e2e"] and (d["e2e"]["ms_per_step"], d["e2e"]["resident"]["ms_per_step"])))
    print("per rank", json.dumps(d["per_rank"]))
    print("selfcheck", d["parity_selfcheck"])
    print("boxes model", d["roofline"].get("boxes_model_ms"))
except Exception as e:
    print("no bench line:", e)
PY
