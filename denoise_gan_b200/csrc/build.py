"""Builds libdg_b200.so in-tree with nvcc for sm_100a (no torch headers, no pybind: the boundary
is the plain C ABI in include/dg_b200.h).  Object files are cached by source mtime."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SOURCES = ["api.cu", "comm.cu", "conv_simt.cu", "conv_umma.cu", "conv_umma_wgrad.cu", "pointwise.cu", "dwconv.cu", "fsrgan_block.cu", "conv_tapsum.cu", "frames.cu", "loss_adam.cu", "pairs.cu", "summaries.cu"]
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith(".cuh")) + ["../../include/dg_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
LIB = os.path.join(PKG, "libdg_b200.so")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    hdrs = [os.path.join(HERE, h) for h in HEADERS]
    nvcc = _nvcc()
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(HERE, s)
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append((s, [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(os.path.join(HERE, "build", name + ".ptxas.log"), "w") as f:
            f.write(r.stderr)
        return name

    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
