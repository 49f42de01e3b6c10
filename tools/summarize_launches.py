"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (one bench run of `--steps 2
--warmup 1`): the launches of the LAST step (between the last two r2c_z launches' positions) as a markdown table.
usage: python tools/summarize_launches.py profiles/launches_<tag>.csv > profiles/launches_<tag>_summary.md"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
L = []
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0}.get(r[iu], 1e-6)
    L.append((r[ik], v))
starts = [i for i, (k, _) in enumerate(L) if k.startswith("r2c_z_kernel") or "r2c_z_kernel" in k]
# the timed step of `--steps 2 --warmup 1` is the last-but-one start .. last start (the last step is followed by draw_qso)
a, b = (starts[-2], starts[-1]) if len(starts) >= 2 else (0, len(L))
tot = OrderedDict()
for k, v in L[a:b]:
    k = k.split("(")[0].replace("void ", "").replace("smk::", "")
    n, t = tot.get(k, (0, 0.0))
    tot[k] = (n + 1, t + v)
step = sum(t for _, t in tot.values())
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (k[:90], n, t, 100 * t / step))
print("| **step** | %d | %.3f | |" % (sum(n for n, _ in tot.values()), step))
rest = [(k.split("(")[0].replace("void ", ""), v) for k, v in L[b:]]
qso = [v for k, v in rest if "draw_qso" in k and "lut" not in k]
if qso:
    print("\n`draw_qso` kernel launches after the steps (outside the step): " + ", ".join("%.3f ms" % v for v in qso))
