"""In-tree build of libbrickscore.so (sm_100a only) with nvcc.

``python -m brickbrain_rec_engine_b200.build`` or ``build()`` from
``__graft_entry__``.  nvcc cross-compiles without a GPU; the shared object is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libbrickscore.so")
# (the tcgen05 scan is one template in bb_scan_tc.cuh; its variants are separate translation units so they compile in parallel)
SOURCES = ["bb_scan_tc_bf16.cu", "bb_scan_tc_split.cu", "bb_scan_tc_pilot.cu", "bb_scan_tc_cg1.cu", "bb_api.cu", "bb_aux.cu", "bb_scan_fp32.cu",
           "bb_itemitem.cu", "bb_spmm.cu", "bb_scan_small.cu"]
HEADERS = ["bb_common.cuh", "bb_kernels.h", "bb_scan_tc.cuh", os.path.join("..", "..", "include", "brickscore.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libbrickscore.so cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    jobs = []
    # BB_EPILOGUE_PROFILE=1: compile the scan's epilogue cycle counters in (bb_info.epilogue_cycles); use with --force
    extra = ["-DBB_EPILOGUE_PROFILE=1"] if os.environ.get("BB_EPILOGUE_PROFILE") == "1" else []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
