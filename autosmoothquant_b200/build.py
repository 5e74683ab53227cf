"""In-tree build of the sm_100a C-ABI library (``libasq_b200.so``) with nvcc.

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box as well as on the
B200 box.  The library links cudart statically and resolves ``cuTensorMapEncodeTiled`` through
``cudaGetDriverEntryPoint`` at run time, so it loads (and exports every symbol of
``include/asq.h``) on machines without a driver.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
SOURCES = [PKG_DIR / "csrc" / "asq_kernels.cu", PKG_DIR / "csrc" / "asq_glue.cu"]
HEADERS = [PKG_DIR / "csrc" / "asq_ptx.cuh", REPO_ROOT / "include" / "asq.h"]
LIB_PATH = PKG_DIR / "libasq_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > built for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if it is missing or older than its sources."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, SOURCES)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
