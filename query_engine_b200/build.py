"""Builds libqgpu.so (the C-ABI library of include/qgpu.h) for sm_100a, in-tree.

    python -m query_engine_b200.build [--force] [--verbose]

Plain nvcc, one object per translation unit (compiled in parallel), linked into
query_engine_b200/lib/libqgpu.so.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
OBJ_DIR = PKG / "build"
LIB = LIB_DIR / "libqgpu.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libqgpu has no CPU build")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    LIB_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "qgpu.h"]
    jobs = []
    objs = []
    for src in _sources():
        obj = OBJ_DIR / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *headers]):
            cmd = [nvcc, *ARCH, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        p = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, p.returncode, p.stdout + p.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, rc, out in ex.map(run, jobs):
                if verbose or rc != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {cmd[-3]}")
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-ldl", "-Xlinker", "--no-undefined"]
        cmd, rc, out = run(cmd)
        if rc != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
            raise RuntimeError("link of libqgpu.so failed")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
