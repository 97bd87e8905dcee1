"""Builds libli_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python linearinterpolators.jl_b200/build.py [--force]

The library links only the CUDA runtime (static); it has no torch / Python dependency.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# development: LI_BUILD_TAG=x builds libli_b200_x.so from build_x/ (A/B builds side by side; load with LI_B200_LIB)
_TAG = os.environ.get("LI_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libli_b200" + ("_" + _TAG if _TAG else "") + ".so")

SOURCES = ["li_api.cu", "li_table.cu", "li_tab_apply.cu", "li_sep.cu", "li_sep2d.cu", "li_sep2d_adjs_f32.cu", "li_sep2d_adjs_f64.cu", "li_affine.cu", "li_affine_f32.cu", "li_affine_f64.cu",
           "li_sep2d_multi_f32.cu", "li_sep2d_multi_f64.cu", "li_comm.cu", "li_vec.cu", "li_affine_gather_f32.cu", "li_affine_gather_f64.cu",
           "li_sep3d_f32.cu", "li_sep3d_f64.cu", "li_sep2d_more_a.cu", "li_sep2d_more_b.cu", "li_sep2d_more_c.cu",
           "li_sep2d_more_d.cu", "li_normal.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    # IEEE everywhere: no fast-math, no flush-to-zero, precise div/sqrt (defaults, stated explicitly)
    "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=hidden",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


# extra flags for experiments, e.g. LI_NVCC_EXTRA="-DLI_DEV_FAST -DADS_NT_=128"
NVCC_FLAGS += os.environ.get("LI_NVCC_EXTRA", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libli_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "li_b200.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc, *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, r.returncode, r.stdout

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, rc, out in ex.map(run, jobs):
            if verbose or rc != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {cmd[-3]}")
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link of libli_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
