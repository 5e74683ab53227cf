"""Golden vectors for the offline calibration step, produced by the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE.  Imports ``autosmoothquant/quantize/calibration.py`` from /root/reference and runs its own
``get_act_scales`` (:44-88), ``get_static_decoder_layer_scales`` (:186-244, all four ``collect_*_layer_scales``),
``get_layers_to_ignore`` (:259-279) and ``quantize_activations_fp8`` (:292-338) on the tiny named-alike models of
``oracle/tiny_models.py``.  The reference tokenises a JSON dataset inside those functions; its ``load_dataset`` name is
pointed at an in-memory list of sample ids and the ``tokenizer`` argument maps a sample id to a pre-generated
``input_ids`` batch, so the reference code runs to the letter on synthetic batches.  Nothing of the reference is
edited or copied; only outputs are stored:

    python oracle/gen_golden_calib.py      # writes tests/golden/calib_golden.{npz,json}
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.gen_golden import import_reference  # noqa: E402
from oracle.tiny_models import ARCHITECTURE, TinyLM, calibration_batches  # noqa: E402

OUT_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"
FAMILIES = ("transformers", "llama", "baichuan", "mixtral")
MODEL_TYPE = {arch: kind for kind, arch in ARCHITECTURE.items()}
MODEL_TYPE["LLaMAForCausalLM"] = "llama"


def import_reference_calibration():
    import_reference()  # layers + the exact-integer _CUDA stub
    models = types.ModuleType("autosmoothquant.models")  # the real package imports HF 4.42 model classes; only the
    models._MODEL_TYPE = MODEL_TYPE                      # architecture -> family table is used by calibration.py
    sys.modules["autosmoothquant.models"] = models
    import autosmoothquant.quantize.calibration as ref_calib
    return ref_calib


class _Samples:
    """Stands in for the HF dataset: ``shuffle`` keeps the order, item i is ``{"text": i}``."""

    def __init__(self, n):
        self.n = n

    def shuffle(self, seed=None):
        return self

    def __getitem__(self, i):
        return {"text": i}


def _tokenizer(batches):
    def tok(text, return_tensors="pt", max_length=None, truncation=True):
        return SimpleNamespace(input_ids=batches[text])
    return tok


def main():
    ref = import_reference_calibration()
    batches = calibration_batches()
    ref.load_dataset = lambda *a, **k: _Samples(len(batches))
    tok = _tokenizer(batches)
    arrays, meta = {}, {"generator": "oracle/gen_golden_calib.py", "torch": torch.__version__, "families": {}}
    for kind in FAMILIES:
        model = TinyLM(kind).eval()
        act = ref.get_act_scales(model, tok, "unused", num_samples=len(batches), seq_len=512)
        for name, v in act.items():
            arrays[f"{kind}.act.{name}"] = v.numpy()
        # the reference does NOT raise top_k in the static pass (no _model_preprocess there): with top_k=2 of 4 experts
        # and 144 calibration tokens every expert of the tiny model is reached; recorded so the test can assert it
        layer_scales, act_dict = ref.get_static_decoder_layer_scales(model, tok, "unused", num_samples=len(batches),
                                                                     seq_len=512, model_type=kind)
        meta["families"][kind] = {"layer_scales": layer_scales, "act_dict": {k: dict(v) for k, v in act_dict.items()},
                                  "top_k_after": getattr(getattr(model.model, "layers", [None])[0] if kind == "mixtral" else None,
                                                         "block_sparse_moe", SimpleNamespace(top_k=None)).top_k}
    # ---- get_layers_to_ignore
    model = TinyLM("llama")
    meta["ignore"] = {
        "re:.*lm_head": sorted(ref.get_layers_to_ignore(model, ["re:.*lm_head"])),
        "exact+regex": sorted(ref.get_layers_to_ignore(model, ["model.layers.0.mlp.down_proj", "re:layers\\.1\\.self_attn"])),
    }
    # ---- static FP8 calibration (reference observers, then FP8LinearStatic.from_float on each)
    ref_linear = sys.modules["autosmoothquant.layers.nn.linear"]
    model = TinyLM("llama").eval()
    ref.quantize_activations_fp8(model, tok, "unused", ["re:.*lm_head"], len(batches))
    fp8 = {}
    for name, mod in model.named_modules():
        if isinstance(mod, ref_linear.FP8StaticLinearQuantizer):
            static = ref_linear.FP8LinearStatic.from_float(mod)
            arrays[f"fp8static.{name}.weight"] = static.weight.view(torch.uint8).numpy()
            fp8[name] = {"weight_scale": float(static.weight_scale), "input_scale": float(static.input_scale),
                         "output_scale": None if static.output_scale is None else float(static.output_scale)}
    meta["fp8_static"] = fp8
    meta["lm_head_is_linear"] = isinstance(model.lm_head, torch.nn.Linear)
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT_DIR / "calib_golden.npz", **arrays)
    (OUT_DIR / "calib_golden.json").write_text(json.dumps(meta, indent=1))
    print(f"wrote {len(arrays)} arrays; fp8 observers: {len(fp8)}; families: {list(meta['families'])}")


if __name__ == "__main__":
    main()
