"""In-tree build of libcubeslam_b200.so (nvcc, sm_100a only).

`python -m cube_slam_wu_b200.build` or `build()`; the .so stays in the package directory (git-ignored, but it
travels to the GPU box with the repo snapshot).
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libcubeslam_b200.so")

CU_SOURCES = ["proposal.cu", "capi_detect.cu", "ba.cu", "ba_solve.cu", "observe.cu", "distmap.cu", "lsd.cu", "lbd.cu", "edlines.cu"]
CPP_SOURCES = ["host_plan.cpp"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # FP64 results must be the plain IEEE sequence (parity with the x86 oracle)
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function",
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, "..", "include", "cubeslam_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    extra = ["-DCSB_SCORE_PHASES"] if os.environ.get("CSB_SCORE_PHASES") == "1" else []  # per-phase cycle counters of k_score (tools/score_phases.py)
    if os.environ.get("CSB_DM_PHASES") == "1":
        extra.append("-DCSB_DM_PHASES")  # per-task stage cycles of k_distmap (tools/distmap_phases.py)
    if os.environ.get("CSB_CHOL_DEBUG") == "1":
        extra.append("-DCSB_CHOL_DEBUG")  # k_chol_solve prints its per-phase cycle counts
    extra += os.environ.get("CSB_NVCC_EXTRA", "").split()  # development: -D overrides for A/B builds (tools/pipelined_time.py)
    for src in CU_SOURCES + CPP_SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s" % src)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
