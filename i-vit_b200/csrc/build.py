"""Build libivit_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python i-vit_b200/csrc/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRCS = ["ivit_api.cu", "ivit_ops.cu", "ivit_fast.cu", "ivit_gemm.cu", "ivit_attn.cu", "ivit_attn_tc.cu", "ivit_attn_pipe.cu", "ivit_swin.cu", "ivit_attn_win.cu", "ivit_tvm.cu"]
HDRS = ["ivit_common.cuh", "ivit_internal.h", "ivit_ptx.cuh", "../../include/ivit_b200.h"]
OUT = os.path.join(HERE, "libivit_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SRCS + HDRS + ["build.py"])


def build(force: bool = False, verbose: bool = False, diag: bool = False) -> str:
    """diag: compile the GEMM's IVIT_GEMM_DEBUG diagnostics in (tools/gemm_bench.py experiments; slower production path)."""
    if not (force or stale()):
        return OUT
    flags = FLAGS + (["-DIVIT_GEMM_DIAG"] if diag else []) + os.environ.get("IVIT_NVCC_EXTRA", "").split()   # e.g. -DIVIT_ATTN_TRACE
    objs = []
    procs = []
    for s in SRCS:
        o = os.path.join(HERE, s.replace(".cu", ".o"))
        objs.append(o)
        procs.append((s, subprocess.Popen([NVCC, *flags, "-c", os.path.join(HERE, s), "-o", o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    bad = False
    for s, p in procs:
        out, _ = p.communicate()
        log.append("==== %s\n%s" % (s, out))
        bad |= p.returncode != 0
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write("\n".join(log))
    if bad:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed, see i-vit_b200/csrc/build.log")
    subprocess.check_call([NVCC, "-shared", "-o", OUT, *objs, "-cudart", "static"])
    if verbose:
        print("\n".join(log))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--diag" in sys.argv, verbose="-v" in sys.argv, diag="--diag" in sys.argv))
