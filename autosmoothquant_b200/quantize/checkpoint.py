"""Checkpoint I/O of quantized models in the reference's on-disk format.

The reference writes ``save_pretrained(output_path)`` + ``quant_config.json`` next to it
(``autosmoothquant/examples/smoothquant_model.py:96-99``) and reads the config back with a bare ``json.load``
(``utils/utils.py:35-39``); the weights file holds the modules' ``state_dict()``: int8 / float8_e4m3fn ``weight``, fp32
``bias``, 0-dim fp32 scale buffers ("Checkpoint").  These helpers write and read exactly that pair without
needing a HF model class: a safetensors file (the format ``save_pretrained`` produces) plus the JSON.  ``"fp8"`` is
normalised to ``"fp8_e4m3"`` on save, as the reference's quantize script does before it writes the file (:69-70).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Tuple

import torch
from torch import nn

WEIGHTS_NAME = "model.safetensors"
CONFIG_NAME = "quant_config.json"


def normalise_type(quant_config: Dict[str, str]) -> Dict[str, str]:
    out = dict(quant_config)
    if out.get("type") == "fp8":
        out["type"] = "fp8_e4m3"
    return out


def save_quantized(model_or_state: "nn.Module | Dict[str, torch.Tensor]", output_path, quant_config: Dict[str, str]) -> Path:
    """Write `model.safetensors` (the state dict, tensors made contiguous, on CPU) and `quant_config.json`."""
    from safetensors.torch import save_file

    out = Path(output_path)
    out.mkdir(parents=True, exist_ok=True)
    state = model_or_state.state_dict() if isinstance(model_or_state, nn.Module) else model_or_state
    tensors = {k: v.detach().cpu().contiguous() for k, v in state.items() if v is not None}
    save_file(tensors, str(out / WEIGHTS_NAME), metadata={"format": "pt"})
    (out / CONFIG_NAME).write_text(json.dumps(normalise_type(quant_config), indent=4))
    return out


def parse_quant_config(config_path) -> Dict[str, str]:
    """utils/utils.py:35-39, plus the load-time normalisation the reference's model classes lack (they compare `type`
    literally to "fp8_e4m3", models/llama.py:76)."""
    with open(config_path, "r", encoding="utf-8") as fh:
        return normalise_type(json.load(fh))


def load_quantized(model_path) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    """(state dict on CPU, quant_config) of a directory written by `save_quantized` or by the reference."""
    from safetensors.torch import load_file

    path = Path(model_path)
    return load_file(str(path / WEIGHTS_NAME)), parse_quant_config(path / CONFIG_NAME)
