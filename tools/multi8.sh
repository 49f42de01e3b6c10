#!/bin/bash
# The 8-GPU session (gpurun --gpus 8): sharded parity at 8 (and 4) ranks incl. the 2048 / 2560-point fused exchanges,
# the nominal chunk (BASELINE config 4) with the persistent-x-pass sweep, and config 5 (all 7 chunks).
tag=${1:-r02p}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "w8- or w4-2048" ) > $O/pytest_multi_8gpu_$tag.log 2>&1
grep -E "passed|failed|skipped|^FAILED|^E  " $O/pytest_multi_8gpu_$tag.log | cut -c1-250 | head -20
timeout 600 $TR bench.py --gpus 8 --steps 4 --warmup 3 --x-sms-sweep ${SWEEP:-0,32,48,64,96} > $O/bench_8gpu_$tag.json 2> $O/bench_8gpu_$tag.err
tail -4 $O/bench_8gpu_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("$O/bench_8gpu_$tag.json").read().strip().splitlines()[-1])
    print("N=8 box", d["config"]["box"], "step %.2f ms  boxes %.2f  skewers %.2f  gather %.2f" % (d["ms_per_step"], d["t_boxes_ms"], d["t_skewers_ms"], d["t_gather_ms"]))
    print("value %.3e  e2e %s" % (d["value"], d["e2e"] and (d["e2e"]["ms_per_step"], d["e2e"]["resident"]["ms_per_step"])))
    print("per rank", json.dumps(d["per_rank"]))
    print("selfcheck", d["parity_selfcheck"])
    print("boxes model", d["roofline"].get("boxes_model_ms"))
    print("x_sms sweep", d["x_sms_sweep_boxes_ms"])
except Exception as e:
    print("no bench line:", e)
PY
timeout 500 $TR tools/run_config5.py --box 2560 --out $O/config5_$tag.json > $O/config5_$tag.log 2> $O/config5_$tag.err
tail -3 $O/config5_$tag.err; tail -2 $O/config5_$tag.log | cut -c1-600
