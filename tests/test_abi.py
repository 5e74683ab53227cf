"""The C-ABI library loads without a GPU and exports exactly what include/asq.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from autosmoothquant_b200 import _lib, build

    build.build()  # no-op when the in-tree .so is newer than its sources
    return _lib.load()


def declared_functions():
    text = (ROOT / "include" / "asq.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asq_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from autosmoothquant_b200 import _lib

    assert declared_functions() == sorted(_lib.EXPORTED_SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_functions():
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr), name


def test_no_torch_or_cublas_dependency():
    """The boundary is a plain C ABI: the .so must not link torch, cuBLAS or CUTLASS-built libraries."""
    import subprocess

    from autosmoothquant_b200 import _lib

    out = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for forbidden in ("libtorch", "libc10", "libcublas", "libcudnn", "libpython"):
        assert forbidden not in out, out


def test_version_and_workspace(lib):
    assert lib.asq_version() == 100
    fixed = lib.asq_workspace_bytes(0, 0)  # counters + stream-K region, what the GEMM-only entry points use
    small = lib.asq_workspace_bytes(1, 16)
    big = lib.asq_workspace_bytes(2048, 4096)
    assert big >= fixed + 2048 * 4096 + 2048 * 4 and small > fixed and big % 1024 == 0


def test_argument_validation_happens_before_any_device_work(lib):
    """Bad arguments are reported as ASQ_ERR_INVALID with a message, GPU or not."""
    rc = lib.asq_i8gemm_o32(None, None, None, 4, 4, 24, None, 0, None)  # K not a multiple of 16
    assert rc == -1
    assert b"multiple of 16" in lib.asq_last_error()
    rc = lib.asq_quantize_act(None, 7, None, None, 4, 32, 0, ctypes.c_float(1.0), 0, 0, None)  # bad dtype
    assert rc == -1
    # asq_fp8_linear_cs: the column-scale vector is mandatory, static activation scales are not a mode of it
    buf = (ctypes.c_char * 4096)()
    addr = (ctypes.addressof(buf) + 1023) & ~1023
    args = [addr, 2, addr, None, addr, 2, 4, 16, 32]
    assert lib.asq_fp8_linear_cs(*args, 2, None, None, 0, None, 0, None) == -1 and b"column scales" in lib.asq_last_error()
    rc = lib.asq_fp8_linear_cs(*args, 1, addr, None, 0, None, 0, None)  # ASQ_ACT_SCALE
    assert rc == -2 and b"dynamic activation scales only" in lib.asq_last_error()  # ASQ_ERR_UNSUPPORTED


def test_fails_loudly_without_a_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    buf = (ctypes.c_char * 4096)()
    addr = (ctypes.addressof(buf) + 15) & ~15
    rc = lib.asq_i8gemm_o32(addr, addr, addr, 16, 16, 16, None, 0, None)
    assert rc == -3  # ASQ_ERR_CUDA: no fallback
    assert lib.asq_device_supported() == 0


def test_header_is_plain_c_and_links(tmp_path):
    """include/asq.h must be consumable from C (the cgo / JNI / ctypes side of the boundary): compile a C99
    translation unit against it with -pedantic and link it with the shared library."""
    import shutil
    import subprocess

    from autosmoothquant_b200 import _lib as L

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = L.LIB_PATH.parent.parent
    src = tmp_path / "t.c"
    src.write_text('#include "asq.h"\n#include <stdio.h>\nint main(void){ printf("%d %zu\\n", asq_version(), '
                   'asq_workspace_bytes(0, 0)); return asq_version() == ASQ_VERSION ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{root / 'include'}", str(src), "-o", str(exe),
                    str(L.LIB_PATH), f"-Wl,-rpath,{L.LIB_PATH.parent}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == L.load().asq_version() and int(out[1]) == L.load().asq_workspace_bytes(0, 0)
