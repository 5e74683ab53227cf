"""Golden vectors for the offline smoothing step, produced by the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE.  Imports ``autosmoothquant/quantize/smooth.py`` from /root/reference (its HF 4.42 model-class
imports are satisfied by whatever the installed transformers provides; the one class it cannot find is replaced
by a stand-in that is never instantiated) and stores inputs / outputs of ``smooth_ln_fcs`` (smooth.py:10-40) for
an OPT-shaped LayerNorm + q/k/v and fc1 pair and a Llama-shaped RMSNorm + gate/up pair:

    python oracle/gen_golden_smooth.py      # writes tests/golden/smooth_golden.npz
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch
from torch import nn

REFERENCE_ROOT = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "smooth_golden.npz"


def import_reference_smooth():
    sys.path.insert(0, str(REFERENCE_ROOT))
    # the Baichuan third-party model file pulls xformers / accelerate-era HF internals: only its class NAMES are
    # needed by smooth.py's isinstance checks, so provide them without importing that file
    bc = types.ModuleType("autosmoothquant.thirdparty.baichuan.modeling_baichuan")

    class RMSNorm(nn.Module):
        pass

    class BaichuanLayer(nn.Module):
        pass

    bc.RMSNorm, bc.BaichuanLayer = RMSNorm, BaichuanLayer
    for name in ("autosmoothquant.thirdparty", "autosmoothquant.thirdparty.baichuan"):
        pkg = types.ModuleType(name)
        pkg.__path__ = []
        sys.modules[name] = pkg
    sys.modules["autosmoothquant.thirdparty.baichuan.modeling_baichuan"] = bc
    root = types.ModuleType("autosmoothquant")
    root.__path__ = [str(REFERENCE_ROOT / "autosmoothquant")]
    sys.modules["autosmoothquant"] = root
    q = types.ModuleType("autosmoothquant.quantize")
    q.__path__ = [str(REFERENCE_ROOT / "autosmoothquant" / "quantize")]
    sys.modules["autosmoothquant.quantize"] = q
    import autosmoothquant.quantize.smooth as ref_smooth  # noqa: E402
    return ref_smooth


def main():
    ref = import_reference_smooth()
    from transformers.models.llama.modeling_llama import LlamaRMSNorm

    g = torch.Generator().manual_seed(0)
    out = {}

    def case(tag, ln, fcs, model_type, alpha):
        act = torch.rand(ln.weight.numel(), generator=g) * 8 + 0.01
        act[::37] *= 40.0  # outlier channels, the SmoothQuant regime
        out[f"{tag}.act_scales"] = act.numpy().copy()
        out[f"{tag}.alpha"] = np.float64(alpha)
        out[f"{tag}.ln_weight_in"] = ln.weight.detach().numpy().copy()
        if getattr(ln, "bias", None) is not None:
            out[f"{tag}.ln_bias_in"] = ln.bias.detach().numpy().copy()
        for i, fc in enumerate(fcs):
            out[f"{tag}.fc{i}_in"] = fc.weight.detach().numpy().copy()
        ref.smooth_ln_fcs(ln, fcs if len(fcs) > 1 else fcs[0], act, model_type, alpha)
        out[f"{tag}.ln_weight_out"] = ln.weight.detach().numpy().copy()
        if getattr(ln, "bias", None) is not None:
            out[f"{tag}.ln_bias_out"] = ln.bias.detach().numpy().copy()
        for i, fc in enumerate(fcs):
            out[f"{tag}.fc{i}_out"] = fc.weight.detach().numpy().copy()

    def lin(i, o):
        m = nn.Linear(i, o, bias=True)
        with torch.no_grad():
            m.weight.copy_(torch.randn(o, i, generator=g) * 0.02)
        return m

    def norm_init(n):
        with torch.no_grad():
            n.weight.copy_(torch.rand(n.weight.numel(), generator=g) + 0.5)
            if getattr(n, "bias", None) is not None:
                n.bias.copy_(torch.randn(n.bias.numel(), generator=g) * 0.1)
        return n

    # OPT-125M proportions (ffn = 4 x hidden) at 1/6 scale so the fixture stays small; the full-size model runs
    # through the same code in tests/test_offline_pipeline.py
    case("opt_qkv", norm_init(nn.LayerNorm(128)), [lin(128, 128) for _ in range(3)], "transformers", 0.5)
    case("opt_fc1", norm_init(nn.LayerNorm(128)), [lin(128, 512)], "transformers", 0.5)
    case("llama_gateup", norm_init(LlamaRMSNorm(96)), [lin(96, 256), lin(96, 256)], "llama", 0.85)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({len(out)} arrays, {OUT.stat().st_size / 1e6:.1f} MB)")


if __name__ == "__main__":
    main()
