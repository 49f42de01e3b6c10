"""2-GPU micro-benchmark of the fused exchange x pass for long lines: torchrun --nproc-per-node 2 tools/bench_p2p.py NX NY NZ"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
from saclaymocks_b200.chunk import ChunkPipeline
NX, NY, NZ = (int(a) for a in sys.argv[1:4])
for mode in ("1", "0"):
    os.environ["SMK_P2P"] = mode
    pipe = ChunkPipeline(NX, NY, NZ, 2.19, device=dev, rank=rank, nranks=world)
    pipe.set_weights({k: torch.ones((NX, NY // world, NZ // 2 + 1), device=dev) for k in ("Pln1", "Pln2", "Pln3", "P0")})
    for it in range(3):
        if it == 1:
            pipe.bs.timing_enable(True)
            torch.cuda.synchronize(); dist.barrier()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
        pipe.step_boxes(seed=it)
    e1.record(); torch.cuda.synchronize()
    t = pipe.bs.timing_collect()
    if rank == 0:
        print("p2p=" + mode, "boxes ms/step %.2f" % (e0.elapsed_time(e1) / 2), {k: round(ms / n, 3) for k, (ms, n) in t.items() if n})
    del pipe
    torch.cuda.empty_cache()
dist.destroy_process_group()
