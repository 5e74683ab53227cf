"""Calibration statistics for SmoothQuant and for static FP8 scales.

What the reference computes in ``autosmoothquant/quantize/calibration.py``:
  * ``get_act_scales`` (:44-88): per-input-channel absmax of every ``nn.Linear`` input -> smoothing factors;
  * ``get_static_decoder_layer_scales`` (:186-244) + the per-family collectors (:90-183): per-tensor absmax of every
    Linear's input and output, regrouped into one dict of ``absmax / 127`` scales per decoder layer under the names
    the family's quantized model class consumes (OPT, Llama, Baichuan — fused ``W_pack`` — and Mixtral, whose experts
    contribute one ``down_input_scales`` entry each);
  * ``quantize_activations_fp8`` (:292-338): every ``nn.Linear`` not excluded by a pattern becomes an
    ``FP8StaticLinearQuantizer`` observer (e4m3 weight, running max of the dynamic per-tensor input scale), the
    calibration batches are run through the model.

The reference tokenises a JSON dataset inside these functions; here the caller passes an iterable of ready
``input_ids`` batches, so nothing depends on a dataset or tokenizer.  Everything is offline torch code and runs on
whatever device the model lives on (CPU for BASELINE config 0, a GPU for real checkpoints).
"""
from __future__ import annotations

import re
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
from torch import nn


class _LinearTap:
    """Forward hooks on every ``nn.Linear`` of a model, feeding (name, input, output) to one reducer."""

    def __init__(self, model: nn.Module, reducer: Callable[[str, torch.Tensor, torch.Tensor], None]):
        self.model, self.reducer, self.handles = model, reducer, []

    def __enter__(self):
        for name, mod in self.model.named_modules():
            if isinstance(mod, nn.Linear):
                self.handles.append(mod.register_forward_hook(self._make(name)))
        return self

    def _make(self, name: str):
        def hook(_module, args, output):
            x = args[0] if isinstance(args, tuple) else args
            y = output[0] if isinstance(output, tuple) else output
            self.reducer(name, x.detach(), y.detach())
        return hook

    def __exit__(self, *exc):
        for h in self.handles:
            h.remove()
        return False


def _run(model: nn.Module, batches: Iterable[torch.Tensor]) -> None:
    device = next(model.parameters()).device
    for input_ids in batches:
        model(input_ids.to(device))


class _AllExpertsRouted:
    """While the smoothing statistics are gathered, a Mixtral router sends every token to EVERY expert, so each
    expert's w1 / w3 sees the whole activation distribution (calibration.py:23-42 raises ``top_k`` of every
    ``block_sparse_moe`` to ``num_local_experts`` and restores it).  No-op for models without such blocks."""

    def __init__(self, model: nn.Module):
        self.blocks = [m for m in model.modules() if hasattr(m, "top_k") and hasattr(m, "experts") and hasattr(m, "gate")]
        self.saved: List[int] = []

    def __enter__(self):
        self.saved = [b.top_k for b in self.blocks]
        for b in self.blocks:
            b.top_k = len(b.experts)
        return self

    def __exit__(self, *exc):
        for b, k in zip(self.blocks, self.saved):
            b.top_k = k
        return False


@torch.no_grad()
def get_act_scales(model: nn.Module, batches: Iterable[torch.Tensor]) -> Dict[str, torch.Tensor]:
    """name -> fp32 CPU vector [in_features]: the largest |x| every input channel of that Linear saw (calibration.py:44-88)."""
    model.eval()
    channel_absmax: Dict[str, torch.Tensor] = {}

    def reduce(name, x, _y):
        seen = x.reshape(-1, x.shape[-1]).abs().amax(dim=0).float().cpu()
        prev = channel_absmax.get(name)
        channel_absmax[name] = seen if prev is None else torch.maximum(prev, seen)

    with _AllExpertsRouted(model), _LinearTap(model, reduce):
        _run(model, batches)
    return channel_absmax


# family -> (prefix of a decoder layer's module name, {scale name: (linear name relative to the layer, "input" | "output")})
_LAYER_KEYS: Dict[str, Tuple[str, Dict[str, Tuple[str, str]]]] = {
    "transformers": ("model.decoder.layers", {          # collect_transformers_layer_scales, calibration.py:90-111
        "attn_input_scale": ("self_attn.q_proj", "input"), "q_output_scale": ("self_attn.q_proj", "output"),
        "k_output_scale": ("self_attn.k_proj", "output"), "v_output_scale": ("self_attn.v_proj", "output"),
        "out_input_scale": ("self_attn.out_proj", "input"), "fc1_input_scale": ("fc1", "input"),
        "fc2_input_scale": ("fc2", "input")}),
    "llama": ("model.layers", {                          # collect_llama_layer_scales, :114-136
        "attn_input_scale": ("self_attn.q_proj", "input"), "q_output_scale": ("self_attn.q_proj", "output"),
        "k_output_scale": ("self_attn.k_proj", "output"), "v_output_scale": ("self_attn.v_proj", "output"),
        "out_input_scale": ("self_attn.o_proj", "input"), "gate_input_scale": ("mlp.gate_proj", "input"),
        "down_input_scale": ("mlp.down_proj", "input")}),
    "baichuan": ("model.layers", {                       # collect_baichuan_layer_scales, :138-156
        "attn_input_scale": ("self_attn.W_pack", "input"), "attn_output_scale": ("self_attn.W_pack", "output"),
        "out_input_scale": ("self_attn.o_proj", "input"), "gate_input_scale": ("mlp.gate_proj", "input"),
        "down_input_scale": ("mlp.down_proj", "input")}),
    "mixtral": ("model.layers", {                        # collect_mixtral_layer_scales, :158-183 (+ the per-expert list below)
        "attn_input_scale": ("self_attn.q_proj", "input"), "q_output_scale": ("self_attn.q_proj", "output"),
        "k_output_scale": ("self_attn.k_proj", "output"), "v_output_scale": ("self_attn.v_proj", "output"),
        "out_input_scale": ("self_attn.o_proj", "input"), "moe_input_scale": ("block_sparse_moe.gate", "input")}),
}


def collect_layer_scales(act_dict: Dict[str, Dict[str, float]], model_type: str, num_layers: int,
                         num_local_experts: int = 0) -> List[Dict[str, object]]:
    """Per-layer scale dicts (absmax / 127) from the per-Linear statistics.  Mixtral adds ``down_input_scales``: one
    entry per expert, from the input of ``block_sparse_moe.experts.<i>.w2`` — an expert no calibration token was routed
    to has no statistics and raises, as in the reference."""
    if model_type not in _LAYER_KEYS:
        raise ValueError(f"unsupport model type: {model_type}")
    prefix, keys = _LAYER_KEYS[model_type]
    out: List[Dict[str, object]] = []
    for idx in range(num_layers):
        scales: Dict[str, object] = {name: act_dict[f"{prefix}.{idx}.{lin}"][io] / 127 for name, (lin, io) in keys.items()}
        if model_type == "mixtral":
            scales["down_input_scales"] = [act_dict[f"{prefix}.{idx}.block_sparse_moe.experts.{e}.w2"]["input"] / 127
                                           for e in range(num_local_experts)]
        out.append(scales)
    return out


@torch.no_grad()
def get_static_decoder_layer_scales(model: nn.Module, batches: Iterable[torch.Tensor], model_type: str = "transformers",
                                    num_layers: Optional[int] = None) -> Tuple[List[Dict[str, object]], Dict[str, Dict[str, float]]]:
    """(per-layer scale dicts, raw per-Linear {"input": absmax, "output": absmax}) — calibration.py:186-244."""
    if model_type not in _LAYER_KEYS:
        raise ValueError(f"unsupport model type: {model_type}")
    model.eval()
    act_dict: Dict[str, Dict[str, float]] = {}

    def reduce(name, x, y):
        stats = act_dict.setdefault(name, {})
        for key, tensor in (("input", x), ("output", y)):
            peak = float(tensor.abs().max())
            stats[key] = max(stats.get(key, peak), peak)

    with _LinearTap(model, reduce):
        _run(model, batches)
    cfg = getattr(model, "config", None)
    if num_layers is None:
        num_layers = cfg.num_hidden_layers
    experts = int(getattr(cfg, "num_local_experts", 0) or 0) if model_type == "mixtral" else 0
    return collect_layer_scales(act_dict, model_type, num_layers, experts), act_dict


# ------------------------------------------------------------------------------------------ FP8 static calibration
def get_layers_to_ignore(model: nn.Module, ignore_patterns: Sequence[str]) -> List[str]:
    """Names of the ``nn.Linear`` modules excluded from FP8 conversion: exact names, or ``re:<regex>`` searched in the
    name (calibration.py:259-279; the quantize script passes ["re:.*lm_head"])."""
    ignored = []
    for name, mod in model.named_modules():
        if not isinstance(mod, nn.Linear):
            continue
        for pat in ignore_patterns:
            hit = re.search(pat[3:], name) is not None if pat.startswith("re:") else pat == name
            if hit:
                ignored.append(name)
                break
    return ignored


def replace_module(model: nn.Module, name: str, new_module: nn.Module) -> None:
    parent_name, _, child = name.rpartition(".")
    setattr(model.get_submodule(parent_name) if parent_name else model, child, new_module)


@torch.no_grad()
def quantize_activations_fp8(model: nn.Module, batches: Iterable[torch.Tensor], ignore_patterns: Sequence[str] = ("re:.*lm_head",),
                             quantize_output: bool = False) -> int:
    """Static FP8 calibration (calibration.py:292-338): every eligible ``nn.Linear`` is replaced IN PLACE by an
    ``FP8StaticLinearQuantizer`` holding its e4m3 weight + per-tensor weight scale; running the batches records the
    largest dynamic per-tensor input scale each observer saw.  ``convert.quantize_linears_fp8`` then turns the observers
    into ``FP8LinearStatic`` modules.  Returns the number of observers installed."""
    from ..layers.functional.quantization import per_tensor_quantize_fp8
    from ..layers.nn.linear import FP8StaticLinearQuantizer

    skip = set(get_layers_to_ignore(model, ignore_patterns))
    targets = [(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear) and n not in skip]
    for name, lin in targets:
        wq, w_scale = per_tensor_quantize_fp8(lin.weight.data)
        bias = lin.bias.data.clone() if lin.bias is not None else None
        replace_module(model, name, FP8StaticLinearQuantizer(lin.in_features, lin.out_features, wq, w_scale, bias,
                                                             quantize_output=quantize_output))
    model.eval()
    _run(model, batches)
    return len(targets)
