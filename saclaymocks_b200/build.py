"""Build libsmk.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsmk.so")
SOURCES = ["smk_boxes.cu", "smk_skewers.cu", "smk_spectra1d.cu", "smk_pk.cu", "smk_qso.cu", "smk_capi.cu"]
# smk_qso.cu restates float64 numpy arithmetic operation by operation: no fused multiply-add contraction
FILE_FLAGS = {"smk_qso.cu": ["-fmad=false"]}
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "smk.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src[:-3] + ".o")
        objs.append(obj)
        cmd = ([nvcc] + flags + FILE_FLAGS.get(src, []) + (["-Xptxas", "-v"] if verbose else [])
               + ["-c", os.path.join(CSRC, src), "-o", obj])
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out = p.communicate()[0].decode()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"])
    return LIB


def build_variant(tag, defines):
    """Kernel-tuning helper: libsmk_<tag>.so compiled with extra -D flags (select it with SMK_LIB_PATH)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")] + ["-D" + d for d in defines]
    out = os.path.join(HERE, "libsmk_%s.so" % tag)
    objs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, "%s_%s.o" % (src[:-3], tag))
        subprocess.check_call([nvcc] + flags + FILE_FLAGS.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj])
        objs.append(obj)
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart"])
    return out


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
