"""Builds ``exahype_b200/libexahype_cuda.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m exahype_b200.build [--force] [--verbose]

One translation unit per physics family, compiled in parallel; ``-fmad=false`` keeps the arithmetic of
``physics.cuh`` contraction-free (bit-exactness against the reference, see csrc/physics.cuh); ``-lineinfo`` so
that ncu's source page maps to these files.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libexahype_cuda.so")

SOURCES = ["exahype_cuda.cu", "inst_euler3d.cu", "inst_euler2d.cu", "inst_swe2d.cu", "inst_fast.cu", "synthetic.cu", "peer_reduce.cu"]
# per-source flags appended after NVCC_FLAGS: the opt-in fast-arithmetic instantiations are the one unit whose
# multiply-add pairs may contract (every kernel in it carries the ArithFast policy type, csrc/physics.cuh)
# (EXAHYPE_FAST_FMAD=false in the environment: tuning builds that keep the branch-free reciprocal / root but not the
# contraction -- what the range checks and slow paths of the IEEE operations cost on their own)
SOURCE_FLAGS = {"inst_fast.cu": ["-fmad=" + os.environ.get("EXAHYPE_FAST_FMAD", "true")]}
HEADERS = ["fv_patch_kernel.cuh", "peer_mail.cuh", "fv3d_march_kernel.cuh", "fv3d_pair_kernel.cuh", "fv2d_march_kernel.cuh", "physics.cuh", "fv_registry.h", os.path.join("..", "..", "include", "exahype_cuda.h")]

NVCC_FLAGS = ["-std=c++17", "-O3", "-fmad=false", "-lineinfo",
              "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def nvcc() -> str:
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found: libexahype_cuda.so cannot be built (there is no CPU fallback)")
    return path


def _host_compiler_args():
    # the image's $CXX (/opt/gcc) works for nvcc; prefer the system g++ when present for a predictable ABI
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, extra_flags=(), lib: str = LIB, build_dir: str = BUILD) -> str:
    """``lib`` / ``build_dir`` / ``extra_flags`` exist for tuning experiments (a second library next to the product one)."""
    global BUILD, LIB
    saved = (BUILD, LIB)
    BUILD, LIB = build_dir, lib
    try:
        return _build(force, verbose, extra_flags)
    finally:
        BUILD, LIB = saved


def _build(force: bool, verbose: bool, extra_flags) -> str:
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for src in SOURCES:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        flags = [f for f in NVCC_FLAGS if not (src in SOURCE_FLAGS and f.startswith("-fmad"))] + SOURCE_FLAGS.get(src, [])
        cmd = [nvcc()] + flags + _host_compiler_args() + list(extra_flags) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as pool:
        for log in pool.map(compile_one, jobs):
            if verbose and log:
                print(log)

    objs = [os.path.join(BUILD, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + _host_compiler_args() + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--variant", default="", help="tuning experiment: build exahype_b200/variants/<name>/libexahype_cuda.so")
    ap.add_argument("--flag", action="append", default=[], help="extra nvcc flag (repeatable), e.g. --flag=-DEXAHYPE_2D_PF=1")
    a = ap.parse_args()
    if a.variant:
        d = os.path.join(HERE, "variants", a.variant)
        print(build(force=a.force, verbose=a.verbose, extra_flags=a.flag, lib=os.path.join(d, "libexahype_cuda.so"),
                    build_dir=os.path.join(d, "build")))
    else:
        print(build(force=a.force, verbose=a.verbose, extra_flags=a.flag))
