"""Host-side behaviour of the reference-API modules (no GPU): constructor contract, checkpoint schema,
buffer placement, from_float arithmetic, quant_config handling and the 'no CPU fallback' rule."""
import numpy as np
import pytest
import torch

from autosmoothquant_b200 import harness
from autosmoothquant_b200.layers.functional import quantization as FQ
from autosmoothquant_b200.layers.nn import linear as NN
from oracle import w8a8_oracle as O


def test_state_dict_schema_matches_reference():
    """Buffer names / dtypes / shapes are the on-disk format (reference linear.py:49-66,137-149,252-256,387-405,515-541)."""
    m = NN.W8A8BFP32OFP32Linear(32, 16, use_bias=True, act_quant="per-token")
    sd = m.state_dict()
    assert sorted(sd) == ["bias", "dequant_scale", "weight"]
    assert sd["weight"].dtype == torch.int8 and tuple(sd["weight"].shape) == (16, 32)
    assert sd["bias"].dtype == torch.float32 and sd["dequant_scale"].dim() == 0
    assert sorted(NN.W8A8BFP32OFP32LinearWithQuantScale(32, 16, act_quant="per-tensor").state_dict()) == \
        ["dequant_scale", "quant_scale", "weight"]
    assert sorted(NN.W8A8BFP32OFP32LinearWithQuantScale(32, 16, act_quant="per-token").state_dict()) == \
        ["dequant_scale", "weight"]
    assert sorted(NN.W8A8BFP32OFP32QKVLinear([8, 4, 4], 32, 16).state_dict()) == \
        ["k_dequant_scale", "q_dequant_scale", "v_dequant_scale", "weight"]
    fp8 = NN.FP8LinearDynamic(32, 16, "per-token", use_bias=True).state_dict()
    assert sorted(fp8) == ["bias", "weight", "weight_scale"] and fp8["weight"].dtype == torch.float8_e4m3fn
    assert sorted(NN.FP8LinearStatic(32, 16).state_dict()) == ["input_scale", "output_scale", "weight", "weight_scale"]
    assert NN.FP8E5M2Linear(32, 16).state_dict()["weight"].dtype == torch.float8_e5m2


def test_act_quant_is_validated():
    with pytest.raises(AssertionError):
        NN.W8A8BFP32OFP32Linear(32, 16, act_quant="per-channel")
    with pytest.raises(AssertionError):
        NN.W8A8BFP32OFP32Linear.from_float(torch.nn.Linear(32, 16), act_quant="nope")


def test_scales_stay_on_host_and_bias_stays_fp32():
    m = NN.W8A8BFP32OFP32LinearWithQuantScale.from_float(torch.nn.Linear(32, 16), 0.05, act_quant="per-tensor")
    m.half()
    assert m.bias.dtype == torch.float32 and m.dequant_scale.device.type == "cpu" and m.quant_scale.device.type == "cpu"
    m.to(torch.bfloat16)
    assert m.dequant_scale.dtype == torch.float32 and m.weight.dtype == torch.int8


def test_from_float_matches_oracle_weight_quantiser():
    torch.manual_seed(1)
    lin = torch.nn.Linear(48, 24)
    w = lin.weight.detach().numpy().copy()
    q, scale = O.quantize_per_tensor_absmax(w, "f32")
    for cls, kw in ((NN.W8A8BFP32OFP32Linear, {}), (NN.W8A8BFP32OFP32LinearWithQuantScale, {})):
        for act in ("per-tensor", "per-token"):
            m = cls.from_float(lin, 0.04, act_quant=act, **kw)
            np.testing.assert_array_equal(m.weight.numpy(), q)
            want = np.float32(scale) if act == "per-token" else np.float32(np.float32(0.04) * np.float32(scale))
            assert abs(float(m.dequant_scale) - float(want)) <= 1e-9
            np.testing.assert_array_equal(m.bias.numpy(), lin.bias.detach().numpy())
    np.testing.assert_array_equal(lin.weight.detach().numpy(), w)  # source module is left untouched
    qkv = NN.W8A8BFP32OFP32QKVLinear.from_float(lin, 0.04, [8, 8, 8], act_quant="per-tensor")
    for i, name in enumerate(("q_dequant_scale", "k_dequant_scale", "v_dequant_scale")):
        qb, sb = O.quantize_per_tensor_absmax(w[8 * i:8 * i + 8], "f32")
        np.testing.assert_array_equal(qkv.weight.numpy()[8 * i:8 * i + 8], qb)
        assert abs(float(getattr(qkv, name)) - float(sb) * 0.04) < 1e-9


def test_fp8_weight_quantiser_matches_oracle():
    torch.manual_seed(2)
    w = torch.randn(24, 64) * 0.05
    q, s = FQ.per_tensor_quantize_fp8(w)
    qo, so = O.quantize_act_fp8(w.numpy(), "f32", "per-tensor")
    np.testing.assert_array_equal(q.view(torch.uint8).numpy(), qo)
    assert float(s) == float(so)


def test_forward_refuses_cpu_tensors():
    m = NN.W8A8BFP32OFP32Linear.from_float(torch.nn.Linear(32, 16), 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(4, 32))
    f = NN.FP8LinearDynamic(32, 16, "per-token")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f(torch.randn(4, 32))


def test_int8gemm_singleton_and_import_aliases():
    import autosmoothquant_b200 as A

    assert NN.Int8GEMM() is NN.Int8GEMM()
    A.install_as_autosmoothquant()
    from autosmoothquant._CUDA import I8CUGEMM  # noqa: F401  (what the reference's linear.py:14 imports)
    from autosmoothquant.layers.nn.linear import W8A8BFP32OFP32Linear

    assert W8A8BFP32OFP32Linear is NN.W8A8BFP32OFP32Linear
    for name in ("linear_a8_w8_o32", "linear_a8_w8_o32_", "linear_a8_w8_o8", "linear_a8_w8_o8_", "linear_a8_w8_b8_o8_"):
        assert callable(getattr(I8CUGEMM(), name))  # the five methods of bindings.cpp:145-155


def test_quant_config_semantics():
    cfg = harness.normalise_quant_config({"type": "fp8"})
    assert cfg["type"] == "fp8_e4m3" and cfg["activation_scheme"] == "dynamic"
    assert harness.normalise_quant_config({})["qkv"] == "per-tensor"
    with pytest.raises(ValueError):
        harness.normalise_quant_config({"qkv": "per-channel"})
    with pytest.raises(ValueError):
        harness.normalise_quant_config({"type": "int4"})
    assert harness.LLAMA2_7B.linear_macs_per_token_per_layer() == 4 * 4096 ** 2 + 3 * 4096 * 11008


def test_fp8_modules_survive_dtype_moves():
    """module.half() / .to(bfloat16) must leave the e4m3 weight, the fp32 bias and the host-side fp32 scales intact
    (float8 is a floating-point dtype to nn.Module._apply); FP8E5M2Linear runs as a plain torch product."""
    import torch
    from autosmoothquant_b200.layers.nn.linear import FP8E5M2Linear, FP8LinearDynamic, FP8LinearStatic

    lin = torch.nn.Linear(32, 16, bias=True)
    static = FP8LinearStatic(32, 16, True)
    static.weight = lin.weight.data.to(torch.float8_e4m3fn)  # torch.empty() storage may hold NaN codes, whose sign a 16-bit trip loses
    for mod in (FP8LinearDynamic.from_float(lin, reference_compat=False), static):
        w0 = mod.weight.view(torch.uint8).clone()
        for move in (lambda m: m.half(), lambda m: m.to(torch.bfloat16), lambda m: m.float()):
            mod = move(mod)
            assert mod.weight.dtype == torch.float8_e4m3fn and torch.equal(mod.weight.view(torch.uint8), w0)
            assert mod.bias.dtype == torch.float32
            for name in mod._scale_names:
                buf = getattr(mod, name)
                assert buf.dtype == torch.float32 and buf.device.type == "cpu" and buf.dim() == 0
    e5 = FP8E5M2Linear.from_float(lin)
    x = torch.randn(3, 32)
    want = torch.nn.functional.linear(x.to(torch.float8_e5m2).float(), lin.weight.data.to(torch.float8_e5m2).float(), lin.bias.data)
    assert torch.allclose(e5(x), want, atol=1e-5)


def test_launches_are_kept_stream_ordered(monkeypatch):
    """Two persistent grids must never share the SMs (their CTAs wait on each other): when a launch arrives on another
    stream than the previous one, the new stream waits for the old one first; inside a CUDA-graph capture nothing is
    recorded.  Exercised with stand-in stream objects (no GPU here)."""
    import torch
    from autosmoothquant_b200 import _lib

    class FakeStream:
        def __init__(self, handle):
            self.cuda_stream, self.waited = handle, []

        def wait_stream(self, other):
            self.waited.append(other.cuda_stream)

    a, b, cap = FakeStream(11), FakeStream(22), FakeStream(33)
    state = {"cur": a, "capturing": False}
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: state["cur"])
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: state["capturing"])
    monkeypatch.setattr(_lib, "_last_launch_stream", {})
    dev = torch.device("cuda", 0)
    assert _lib._stream(dev) == 11 and _lib._stream(dev) == 11 and a.waited == []  # same stream: nothing to do
    state["cur"] = b
    assert _lib._stream(dev) == 22 and b.waited == [11]  # b waits for everything queued on a
    assert _lib._stream(dev) == 22 and b.waited == [11]
    state.update(cur=cap, capturing=True)
    assert _lib._stream(dev) == 33 and cap.waited == []  # capture: no cross-stream event, bookkeeping untouched
    state.update(cur=a, capturing=False)
    assert _lib._stream(dev) == 11 and a.waited == [22]
    monkeypatch.setattr(_lib, "_STREAM_GUARD", False)
    state["cur"] = b
    assert _lib._stream(dev) == 22 and b.waited == [11]  # guard off: no new wait
