"""Micro-benchmark of the FFT passes on one GPU for an arbitrary box (kernel tuning)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from saclaymocks_b200.boxes import BoxSynth
NX, NY, NZ = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
bs = BoxSynth(NX, NY, NZ, 2.19, device=dev)
boxk = bs.draw_grf_boxk(seed=1)
out = bs.empty_box()
W = torch.ones((NX, NY, NZ // 2 + 1), dtype=torch.float32, device=dev)
for name, wt in (("eta_xy", None), ("boxln_1", W)):
    for _ in range(2):
        bs.synth(boxk, name, wtable=wt, store_p0=False, out=out)
    bs.timing_enable(True)
    for _ in range(reps):
        bs.synth(boxk, name, wtable=wt, store_p0=False, out=out)
    t = bs.timing_collect()
    bs.timing_enable(False)
    nk = NX * NY * (NZ // 2 + 1) * 8
    mult = {"inv_yz": 4, "fwd_zy": 3}      # chained pairs: bytes of both passes under the 24 B/cell model
    tot = sum(ms / n for ms, n in t.values() if n)
    print(name, "total %.3f ms" % tot,
          {k: "%.3f ms %.0f GB/s" % (ms / n, mult.get(k, 2) * nk / (ms / n * 1e-3) / 1e9) for k, (ms, n) in t.items() if n})
