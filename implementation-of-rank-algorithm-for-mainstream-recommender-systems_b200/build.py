"""Builds librank_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library links only the CUDA runtime: no libtorch, no pybind.  Objects are rebuilt when a
source or header is newer than them; `python -m ...build` or `__graft_entry__.build()` run it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
INCLUDE = REPO_ROOT / "include"
OBJ_DIR = PKG_DIR / "build"
LIB_PATH = PKG_DIR / "librank_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: librank_b200.so cannot be built")


def _newer(src: Path, dst: Path) -> bool:
    return (not dst.exists()) or src.stat().st_mtime > dst.stat().st_mtime


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link librank_b200.so.  Returns its path."""
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    headers = list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    hdr_time = max(h.stat().st_mtime for h in headers)
    jobs = []
    objs = []
    for src in sources():
        obj = OBJ_DIR / (src.stem + ".o")
        objs.append(obj)
        if force or _newer(src, obj) or obj.stat().st_mtime < hdr_time:
            cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-I", str(CSRC), "-c", str(src),
                   "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return res.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    if jobs or not LIB_PATH.exists():
        link = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs), "--cudart", "static"]
        run(link)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
