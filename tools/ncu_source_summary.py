"""Condense an `ncu --page source --csv` export: stall reasons summed over the kernel and the instructions with the most
stall samples.   usage: python tools/ncu_source_summary.py <source.csv> [kernel index] [top n]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    heads = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r]
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    lo = heads[k]
    hi = heads[k + 1] - 1 if k + 1 < len(heads) else len(rows)
    h = rows[lo]
    body = [r for r in rows[lo + 1:hi] if len(r) == len(h)]
    col = {c: i for i, c in enumerate(h)}
    stalls = [c for c in h if c.startswith("stall_")]
    tot = {c: sum(float(r[col[c]] or 0) for r in body) for c in stalls}
    allsum = sum(tot.values())
    print("instructions: %d   stall samples: %d" % (len(body), allsum))
    for c, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print("  %-28s %6.2f %%" % (c, 100 * v / allsum))
    samp = "Warp Stall Sampling (All Samples)"
    ex = "Instructions Executed"
    print("top instructions by stall samples:")
    for r in sorted(body, key=lambda r: -float(r[col[samp]] or 0))[:top]:
        print("  %6s  %5.2f%%  exec %-10s %s" % (r[col["Address"]][-5:], 100 * float(r[col[samp]] or 0) / max(allsum, 1),
                                                  r[col[ex]], r[col["Source"]][:90]))
    wf, ideal = "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"
    if wf in col:
        print("shared wavefronts %.3g (ideal %.3g)" % (sum(float(r[col[wf]] or 0) for r in body),
                                                        sum(float(r[col[ideal]] or 0) for r in body)))
    print("executed warp instructions %.4g" % sum(float(r[col[ex]] or 0) for r in body))


if __name__ == "__main__":
    main()
