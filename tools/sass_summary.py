"""SASS opcode summary per kernel of libsmk.so (cuobjdump -sass): the evidence table kept under profiles/.
usage: python tools/sass_summary.py [libsmk.so] > profiles/sass_summary_<tag>.md"""
import collections
import re
import subprocess
import sys

CLASSES = [("UTMALDG", r"^UTMALDG"), ("UBLKCP", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"), ("LDGSTS", r"^LDGSTS"),
           ("FFMA2/FMUL2/FADD2", r"^(FFMA2|FMUL2|FADD2)"), ("FFMA/FMUL/FADD", r"^(FFMA|FMUL|FADD)(\.|$)"),
           ("DFMA/DMUL/DADD", r"^(DFMA|DMUL|DADD)"), ("MUFU", r"^MUFU"), ("LDS", r"^LDS"), ("STS", r"^STS"),
           ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDG/STG.128", r"^(LDG|STG).*\.128"), ("ATOM/RED", r"^(ATOM|RED|ATOMG|REDG)"),
           ("REDUX/SHFL", r"^(REDUX|SHFL)"), ("HMMA/UTC*MMA", r"^(HMMA|UTC.*MMA)"), ("CCTL (prefetch/discard)", r"^CCTL"),
           ("BAR", r"^BAR"), ("LDL/STL (spills)", r"^(LDL|STL)")]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "saclaymocks_b200/libsmk.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for name, pat in CLASSES:
                if re.match(pat, op):
                    kernels[cur][name] += 1
    demangled = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print("| kernel | instructions | " + " | ".join(n for n, _ in CLASSES) + " |")
    print("|---|---|" + "---|" * len(CLASSES))
    for (k, c), d in zip(kernels.items(), demangled):
        name = re.sub(r"\(.*\)$", "", d).replace("void smk::", "").replace("smk::", "")
        print("| `%s` | %d | " % (name[:90], c["total"]) + " | ".join(str(c[n]) if c[n] else "" for n, _ in CLASSES) + " |")


if __name__ == "__main__":
    main()
