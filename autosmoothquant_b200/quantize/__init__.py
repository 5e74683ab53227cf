"""Offline SmoothQuant pipeline (calibrate -> smooth -> quantize -> save): CPU/GPU torch plumbing around the hot path."""
from .calibration import (get_act_scales, get_layers_to_ignore, get_static_decoder_layer_scales,  # noqa: F401
                          quantize_activations_fp8)
from .checkpoint import load_quantized, parse_quant_config, save_quantized  # noqa: F401
from .convert import quantize_decoder_layers, quantize_linears_fp8  # noqa: F401
from .smooth import smooth_lm, smooth_ln_fcs  # noqa: F401
