"""TEST INFRASTRUCTURE — builds the reference's own native INT8 GEMM extension into ``oracle/_ref/``.

The reference's hot-path arithmetic lives in ``csrc/int8gemm`` (a pybind11 class ``I8CUGEMM`` around
cuBLASLt, ``/root/reference/csrc/int8gemm/bindings.cpp:145-155``, built by ``/root/reference/setup.py:6-20``).
It is four host-only C++ files, so it compiles straight from where the sources lie with ``g++`` — no
copy of the sources enters this repository, only the resulting ``_CUDA*.so`` lands in ``oracle/_ref/``
(git-ignored, but it travels to the GPU box).  Differences from the reference's ``setup.py``: cuBLAS /
cuBLASLt are linked dynamically (the static link produces an 808 MB file), nothing else.

The extension needs a GPU at run time (``I8CUGEMM()`` creates a cuBLASLt handle and captures a CUDA
stream), so it is only *built* here; ``tests/test_ref_extension.py`` (``-m gpu``) runs it on the B200 as a
second, independent checker for ``asq_i8gemm_o32`` and the fused modules.  Only tests, ``smoke()`` and
``bench.py``'s reference legs may load it; the product never does.
"""
from __future__ import annotations

import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT_DIR = HERE / "_ref"
REF_SRC = Path("/root/reference/csrc/int8gemm")
SOURCES = ["cuda_utils.cc", "cublasINT8MMWrapper.cc", "cublasAlgoMap.cc", "bindings.cpp"]


def lib_path() -> Path:
    return OUT_DIR / ("_CUDA" + sysconfig.get_config_var("EXT_SUFFIX"))


def available() -> bool:
    return lib_path().exists()


def build(force: bool = False, verbose: bool = False) -> Path | None:
    """Compile the reference extension if its sources are present (they are not on the GPU box)."""
    out = lib_path()
    if not REF_SRC.is_dir():
        return out if out.exists() else None
    if out.exists() and not force:
        newest = max((REF_SRC / s).stat().st_mtime for s in SOURCES)
        if out.stat().st_mtime >= newest:
            return out
    import torch  # noqa: F401  (only for the include / library paths)
    from torch.utils import cpp_extension as ce

    OUT_DIR.mkdir(parents=True, exist_ok=True)
    obj_dir = OUT_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    incs = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    cxx = ["g++", "-std=c++17", "-O3", "-fPIC", "-DTORCH_EXTENSION_NAME=_CUDA", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI)), "-w"]
    objs = []
    procs = []
    for s in SOURCES:
        o = obj_dir / (Path(s).stem + ".o")
        objs.append(str(o))
        cmd = [*cxx, *incs, "-c", str(REF_SRC / s), "-o", str(o)]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"reference extension failed to compile:\n{' '.join(cmd)}\n{log}")
        if verbose and log:
            print(log)
    libdirs = ce.library_paths("cuda")
    link = ["g++", "-shared", "-o", str(out), *objs, *[f"-L{d}" for d in libdirs],
            *[f"-Wl,-rpath,{d}" for d in libdirs],
            "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
            "-lcublas", "-lcublasLt", "-lcudart", "-lrt", "-lpthread", "-ldl"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"reference extension failed to link:\n{' '.join(link)}\n{res.stdout}\n{res.stderr}")
    for o in objs:
        Path(o).unlink(missing_ok=True)
    return out


def load():
    """Import the built extension as a module object (GPU boxes only: construction needs a device)."""
    import importlib.util

    import torch  # noqa: F401  (the extension links against libtorch)

    path = lib_path()
    if not path.exists():
        raise FileNotFoundError(f"{path} not built (run oracle/build_ref.py where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("_CUDA", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
