"""Pin the numpy oracle against golden vectors recorded from the unmodified reference (CPU).

The reference has no tests or fixtures of its own (its tests/__init__.py is empty); the vectors
were produced by oracle/gen_golden.py, which imports /root/reference's Linear classes and
quantisation helpers with a stub integer-GEMM extension.  Integer results must match exactly,
16/32-bit float outputs bit-for-bit (the oracle restates every rounding step).
"""
import numpy as np
import pytest

from oracle import w8a8_oracle as O


def _cases(golden, kind):
    out = [c for c in golden if c["kind"] == kind]
    assert out, f"no golden cases of kind {kind}"
    return out


def test_golden_inventory(golden):
    kinds = {c["kind"] for c in golden}
    assert kinds == {"int8_linear", "int8_qkv", "weight_quant", "fp8_per_token", "fp8_static",
                     "fp8_per_tensor", "fp8_linear"}
    assert len(golden) >= 100


def test_int8_linear_bit_exact(golden):
    for c in _cases(golden, "int8_linear"):
        y = O.w8a8_linear(
            c["x"], c["dtype"], c["weight"], c["dequant_scale"], act_quant=c["act_quant"],
            bias=c.get("bias"), quant_scale=c.get("quant_scale"), div_mode="exact",
        )
        np.testing.assert_array_equal(y, c["y"], err_msg=f"{c['id']} {c['cls']} {c['act_quant']} {c['dtype']}")


def test_int8_linear_covers_every_variant(golden):
    seen = {(c["cls"], c["act_quant"], c["dtype"], "bias" in c) for c in _cases(golden, "int8_linear")}
    for cls in ("W8A8BFP32OFP32Linear", "W8A8BFP32OFP32LinearWithQuantScale"):
        for aq in ("per-tensor", "per-token"):
            for dt in ("f32", "f16", "bf16"):
                for b in (False, True):
                    assert (cls, aq, dt, b) in seen


def test_int8_qkv_bit_exact(golden):
    for c in _cases(golden, "int8_qkv"):
        y = O.w8a8_qkv_linear(c["x"], c["dtype"], c["weight"], c["qkv_size"], c["q_scale"], c["k_scale"],
                              c["v_scale"], act_quant=c["act_quant"], bias=c.get("bias"), div_mode="exact")
        np.testing.assert_array_equal(y, c["y"], err_msg=c["id"])


def test_weight_quant_exact(golden):
    for c in _cases(golden, "weight_quant"):
        q, s = O.quantize_per_tensor_absmax(c["w"], c["dtype"])
        assert float(s) == c["scale"], c["id"]
        np.testing.assert_array_equal(q, c["q"], err_msg=c["id"])


@pytest.mark.parametrize("kind,mode", [("fp8_per_token", "per-token"), ("fp8_static", "scale"),
                                       ("fp8_per_tensor", "per-tensor")])
def test_fp8_quantisers_bit_exact(golden, kind, mode):
    for c in _cases(golden, kind):
        q, s = O.quantize_act_fp8(c["x"], c["dtype"], mode, c.get("in_scale", 1.0), div_mode="exact")
        want = c["q"]
        # NaN payloads (0/0 rows) may differ in sign between libraries: compare NaN-ness, then bytes
        nan_w = (want & 0x7F) == 0x7F
        nan_g = (q & 0x7F) == 0x7F
        np.testing.assert_array_equal(nan_g, nan_w, err_msg=c["id"])
        np.testing.assert_array_equal(q[~nan_w], want[~nan_w], err_msg=c["id"])
        if kind == "fp8_per_token":
            np.testing.assert_array_equal(s, c["scale"], err_msg=c["id"])
        elif kind == "fp8_per_tensor":
            assert float(np.asarray(s).reshape(())) == c["scale"], c["id"]


def test_fp8_linear_reference_math(golden):
    """The reference's fp8 forward is dequantise + GEMM in the activation dtype; BLAS summation order is not
    pinned, so the tolerance is 1e-5 of the output scale for fp32 and 2 ulp for bf16.  Output fake-quantisation
    snaps values to the e4m3 grid, so there a handful of values may land on a neighbouring code."""
    seen = set()
    for c in _cases(golden, "fp8_linear"):
        act = c["act"]
        seen.add((act, c["dtype"], "out_scale" in c))
        y = O.fp8_linear_reference_math(c["x"], c["dtype"], c["w"], c["w_scale"], act_quant=act,
                                        in_scale=c.get("in_scale", 1.0), bias=c.get("bias"), div_mode="exact",
                                        out_scale=c.get("out_scale", 0.0))
        scale = np.abs(c["y"]).max()
        if "out_scale" in c:
            assert np.mean(y != c["y"]) < 0.02 and np.abs(y - c["y"]).max() <= 0.13 * scale, c["id"]
            continue
        tol = 1e-5 if c["dtype"] == "f32" else 2 ** -7
        np.testing.assert_allclose(y, c["y"], rtol=0, atol=tol * scale, err_msg=c["id"])
        if c["dtype"] == "f32" and act != "per-tensor":
            # and the fp64 evaluation of the same quantised operands agrees to fp32 accumulation error
            mode = "per-token" if act == "per-token" else "scale"
            q, s = O.quantize_act_fp8(c["x"], "f32", mode, c.get("in_scale", 1.0))
            a_scale = s if mode == "per-token" else np.full(c["x"].shape[0], c["in_scale"], np.float32)
            y64 = O.fp8_linear_exact(q, c["w"], a_scale, c["w_scale"], c.get("bias"))
            np.testing.assert_allclose(y64, c["y"], rtol=0, atol=1e-5 * scale, err_msg=c["id"])
    assert ("per-tensor", "bf16", False) in seen and ("static", "f32", True) in seen
