"""Offline SmoothQuant pipeline (calibrate -> smooth -> quantize): CPU/GPU torch plumbing around the hot path."""
from .calibration import get_act_scales, get_static_decoder_layer_scales  # noqa: F401
from .convert import quantize_decoder_layers  # noqa: F401
from .smooth import smooth_lm, smooth_ln_fcs  # noqa: F401
