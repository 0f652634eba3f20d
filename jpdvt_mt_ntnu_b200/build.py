"""In-tree nvcc build of libjpdvt_sm100.so (sm_100a only; cross-compiles on a box without a GPU).

    python -m jpdvt_mt_ntnu_b200.build [--force] [--verbose]

The shared library is linked against the static CUDA runtime and resolves the driver's tensor-map encoder through
cudaGetDriverEntryPoint, so it loads (and its symbols can be enumerated) on a CPU-only box too.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjpdvt_sm100.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
SOURCES = ["api.cu", "train_api.cu", "gemm.cu", "attention.cu", "attention_tc.cu", "elementwise.cu", "assign.cu", "puzzle.cu", "fold.cu", "backward.cu", "optim.cu", "loss.cu", "attention_bwd_tc.cu", "peer_optim.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libjpdvt_sm100.so cannot be built (there is no CPU fallback)")


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(os.path.join(CSRC, name), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "jpdvt_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    return open(STAMP).read().strip() != _digest()


def _unit_digest(src: str) -> str:
    """Digest of one translation unit: its own text, every shared header and the flags."""
    h = hashlib.sha256()
    h.update(open(os.path.join(CSRC, src), "rb").read())
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cuh", ".h")):
            h.update(open(os.path.join(CSRC, name), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "jpdvt_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(CSRC, src[:-3] + ".o")
        tag = obj + ".sha"
        dig = _unit_digest(src)
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read().strip() == dig:
            return src, obj, subprocess.CompletedProcess([], 0, "", "(up to date)")
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            with open(tag, "w") as f:
                f.write(dig)
        return src, obj, r

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, obj, r in ex.map(compile_one, _sources()):
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(f"== {src}\n{r.stderr}\n")
            objs.append(obj)
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
