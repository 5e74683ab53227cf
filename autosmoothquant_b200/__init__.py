"""B200-native (sm_100a) SmoothQuant W8A8 / FP8 linear path behind the ``autosmoothquant.layers`` API.

Public surface
  autosmoothquant_b200.layers.nn.linear   the reference's quantized Linear module classes
  autosmoothquant_b200._CUDA.I8CUGEMM     the reference's native INT8 GEMM class (5 methods)
  autosmoothquant_b200._lib               ctypes binding of the C ABI in include/asq.h
  autosmoothquant_b200.tp                 column / row tensor-parallel wrappers (NCCL all-reduce)
  install_as_autosmoothquant()            make `import autosmoothquant.layers...` resolve to this package

The compute path is one fused CUDA kernel per Linear call; there is no CPU or eager fallback.
"""
from __future__ import annotations

import sys
import types

__version__ = "0.1.0"


def install_as_autosmoothquant(force: bool = False) -> None:
    """Register this package's modules under the reference's import names.

    After this call ``from autosmoothquant.layers.nn.linear import W8A8BFP32OFP32Linear`` and
    ``from autosmoothquant._CUDA import I8CUGEMM`` (what the reference's model classes import,
    autosmoothquant/models/llama.py:18-21) resolve to the B200 implementations.  If the real
    ``autosmoothquant`` package is importable its ``layers`` and ``_CUDA`` submodules are shadowed,
    the rest (models, quantize, utils) keeps working on top of them.
    """
    from . import _CUDA, layers
    from .layers import functional, nn
    from .layers.functional import quantization
    from .layers.nn import bmm, linear

    if "autosmoothquant" not in sys.modules or force:
        try:
            import autosmoothquant as root  # the reference, if installed
        except Exception:  # noqa: BLE001 - absent or broken install: provide a namespace shell
            root = types.ModuleType("autosmoothquant")
            root.__path__ = []  # type: ignore[attr-defined]
            sys.modules["autosmoothquant"] = root
    root = sys.modules["autosmoothquant"]
    mapping = {
        "autosmoothquant._CUDA": _CUDA,
        "autosmoothquant.layers": layers,
        "autosmoothquant.layers.nn": nn,
        "autosmoothquant.layers.nn.linear": linear,
        "autosmoothquant.layers.nn.bmm": bmm,
        "autosmoothquant.layers.functional": functional,
        "autosmoothquant.layers.functional.quantization": quantization,
    }
    for name, mod in mapping.items():
        sys.modules[name] = mod
    root._CUDA = _CUDA
    root.layers = layers
