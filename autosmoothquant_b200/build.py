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
SOURCES = [PKG_DIR / "csrc" / "asq_kernels.cu", PKG_DIR / "csrc" / "asq_glue.cu", PKG_DIR / "csrc" / "asq_smallm.cu",
           PKG_DIR / "csrc" / "asq_nvls_probe.cu"]
HEADERS = [PKG_DIR / "csrc" / "asq_ptx.cuh", PKG_DIR / "csrc" / "asq_smallm.h", REPO_ROOT / "include" / "asq.h"]
# experiments: ASQ_LIB_NAME=libasq_b200_x.so ASQ_NVCC_DEFS="-DASQ_EPI_NBUF=1" builds (and _lib loads) a variant library
LIB_PATH = PKG_DIR / os.environ.get("ASQ_LIB_NAME", "libasq_b200.so")

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


N_KERNEL_TUS = 24  # asq_kernels.cu is compiled once per kernel instantiation set (-DASQ_TU=1..24) + once for the host side (0)
OBJ_DIR = Path(os.environ.get("ASQ_OBJ_DIR", "/tmp/asq_b200_obj" + ("_" + os.environ["ASQ_LIB_NAME"] if "ASQ_LIB_NAME" in os.environ else "")))  # objects stay out of the tree


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if it is missing or older than its sources.  The translation units (one per
    instantiation of the big kernel, the host side, the glue kernels) are compiled in parallel, then linked."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + ["-c"] + os.environ.get("ASQ_NVCC_DEFS", "").split()
    if verbose:
        compile_flags += ["-Xptxas", "-v"]
    jobs = []
    kernels_cu, glue_cu, smallm_cu, probe_cu = SOURCES
    with_mc = "-DASQ_ENABLE_MC" in compile_flags  # TUs 5 / 6: the 4-CTA multicast experiment (ASQ_MC=2), off by default
    for tu in range(N_KERNEL_TUS + 1):
        if tu in (5, 6) and not with_mc:
            continue
        obj = OBJ_DIR / f"asq_kernels_tu{tu}.o"
        jobs.append((obj, [nvcc, *compile_flags, f"-DASQ_TU={tu}", "-o", str(obj), str(kernels_cu)]))
    obj = OBJ_DIR / "asq_glue.o"
    jobs.append((obj, [nvcc, *compile_flags, "-o", str(obj), str(glue_cu)]))
    obj = OBJ_DIR / "asq_smallm.o"
    jobs.append((obj, [nvcc, *compile_flags, "-o", str(obj), str(smallm_cu)]))
    obj = OBJ_DIR / "asq_nvls_probe.o"
    jobs.append((obj, [nvcc, *compile_flags, "-o", str(obj), str(probe_cu)]))
    procs = [(obj, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for obj, cmd in jobs]
    logs = []
    for obj, cmd, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{out}")
        logs.append(out)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB_PATH),
            *[str(obj) for obj, _ in jobs]]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{' '.join(link)}\n{res.stdout}\n{res.stderr}")
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


CEILING_SRC = PKG_DIR / "csrc" / "asq_ceiling.cu"
CEILING_LIB = PKG_DIR / "libasq_ceiling.so"


def build_ceiling(force: bool = False) -> Path:
    """The tcgen05 issue-rate microbenchmark (csrc/asq_ceiling.cu -> libasq_ceiling.so): a measurement utility with
    its own library, so the drop-in library of include/asq.h is not touched by it (scripts/int8_ceiling.py, bench.py)."""
    deps = [CEILING_SRC, PKG_DIR / "csrc" / "asq_ptx.cuh"]
    if not force and CEILING_LIB.exists() and all(p.stat().st_mtime <= CEILING_LIB.stat().st_mtime for p in deps):
        return CEILING_LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", str(CEILING_LIB), str(CEILING_SRC)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    return CEILING_LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_ceiling(force=True))
