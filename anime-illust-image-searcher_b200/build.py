"""Builds libais_b200.so (the C ABI + every CUDA kernel) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libais_b200.so")
SOURCES = ["engine.cu", "pickle_csr.cpp"]
HEADERS = ["common.cuh", "finals.cuh", "scan.cuh", "scan_tc.cuh", "scan_pair.cuh", "bm25.cuh", "build.cuh", "select2.cuh", "select.cuh", os.path.join("..", "..", "include", "ais_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-Xptxas", "-v"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: ais_b200 needs the CUDA toolkit to build (there is no CPU path)")


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT + ".tmp"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed (%d)" % res.returncode)
    os.replace(OUT + ".tmp", OUT)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
