"""GPU parity against the reference's OWN native code: ``oracle/_ref/_CUDA*.so`` is the reference's
``csrc/int8gemm`` extension (pybind class ``I8CUGEMM``, ``/root/reference/csrc/int8gemm/bindings.cpp:145-155``)
compiled from its unmodified sources by ``oracle/build_ref.py``.  This is the path north_star asks to
match "to <= 1 LSB of the INT32 accumulator"; the bar here is 0 LSB.

The reference hard-codes a cuBLASLt algorithm (id 21, tile 20, stages 17:
``cublasINT8MMWrapper.cc:313-339``) and ignores the matmul status (``:343``).  If cuBLASLt rejects
that algorithm on sm_100 the output buffer is simply left untouched; the tests detect that (sentinel
fill) and report it as an xfail naming the cause instead of comparing garbage.

Nothing here reads /root/reference: the prebuilt .so travels with the repository snapshot.
"""
import numpy as np
import pytest
import torch

from oracle import build_ref
from oracle import w8a8_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from autosmoothquant_b200 import _lib as L
    from autosmoothquant_b200.layers.nn import linear as NN

DEV = torch.device("cuda:0")
SENTINEL = -1234567


@pytest.fixture(scope="module")
def ref_gemm():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    if not build_ref.available():
        pytest.skip("oracle/_ref extension not built (oracle/build_ref.py needs /root/reference)")
    torch.cuda.set_device(0)
    torch.zeros(1, device=DEV)  # the reference captures the current stream at construction
    return build_ref.load().I8CUGEMM()


def ref_o32(ref_gemm, a, w):
    """linear_a8_w8_o32_ exactly as linear.py:97-103 calls it; returns None if the GEMM did not run."""
    out = torch.full((a.shape[0], w.shape[0]), SENTINEL, dtype=torch.int32, device=DEV)
    ref_gemm.linear_a8_w8_o32_(a, w, out)
    torch.cuda.synchronize()
    if bool((out == SENTINEL).all()):
        return None
    return out


def need(out):
    if out is None:
        pytest.xfail("the reference's hard-coded cuBLASLt algo (id 21 / tile 20 / stages 17) is rejected on "
                     "sm_100 and its status is ignored: output untouched, nothing to compare")
    return out


@pytest.mark.parametrize("M,N,K", [(16, 64, 64), (128, 256, 128), (256, 512, 1024), (2048, 4096, 4096),
                                   (2048, 4096, 11008), (333, 1000, 4096)])
def test_i8gemm_o32_equals_reference_extension(ref_gemm, M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).to(DEV)
    w = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g).to(DEV)
    ours = torch.empty((M, N), dtype=torch.int32, device=DEV)
    L.i8gemm_o32(a, w, ours)
    torch.cuda.synchronize()
    if M * N * K <= 256 * 512 * 1024:
        np.testing.assert_array_equal(ours.cpu().numpy(), O.int8_gemm_i32(a.cpu().numpy(), w.cpu().numpy()))
    ref = need(ref_o32(ref_gemm, a, w))
    assert torch.equal(ours, ref), f"int32 accumulators differ from the reference extension at {(M, N, K)}"


def reference_forward(ref_gemm, x, weight, dequant_scale, bias, act_quant, quant_scale=None):
    """The eager launches of W8A8BFP32OFP32Linear(.WithQuantScale).forward (linear.py:83-106, 278-302)
    on the GPU with the reference's native GEMM in the middle."""
    shape, dtype = x.shape, x.dtype
    x2 = x.view(-1, shape[-1])
    if act_quant == "per-token":
        qs = x2.abs().max(dim=-1, keepdim=True)[0].div(127.0).to(torch.float32)
        q = (x2 / qs).round().clamp(-128, 127).to(torch.int8)
        ds = float(dequant_scale) * qs
    elif quant_scale is not None:
        q = (x2 / float(quant_scale)).round().clamp(-128, 127).to(torch.int8)
        ds = float(dequant_scale)
    else:
        q = x2.round().clamp(-128, 127).to(torch.int8)
        ds = float(dequant_scale)
    acc = ref_o32(ref_gemm, q, weight)
    if acc is None:
        return None
    out = ds * acc + bias if bias is not None else ds * acc
    return out.view(*shape[:-1], -1).to(dtype)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("cls,act", [("W8A8BFP32OFP32Linear", "per-tensor"), ("W8A8BFP32OFP32Linear", "per-token"),
                                     ("W8A8BFP32OFP32LinearWithQuantScale", "per-tensor"),
                                     ("W8A8BFP32OFP32LinearWithQuantScale", "per-token")])
def test_fused_module_equals_reference_path_on_gpu(ref_gemm, cls, act, dtype):
    """Same inputs, same device: our one fused launch against the reference's eager prologue + native
    GEMM + eager epilogue.  Default division mode (= torch CUDA's reciprocal multiply).  Rows whose
    per-token scale is 0 are excluded: 0/0 -> NaN -> int8 is undefined behaviour in the eager path."""
    g = torch.Generator().manual_seed(11)
    M, K, N = 512, 1024, 768
    lin = torch.nn.Linear(K, N, bias=True)
    with torch.no_grad():
        lin.weight.copy_(torch.randn(N, K, generator=g) * 0.05)
        lin.bias.copy_(torch.randn(N, generator=g))
    mod = getattr(NN, cls).from_float(lin, 0.04, act_quant=act).to(DEV)
    scale = 30.0 if (cls == "W8A8BFP32OFP32Linear" and act == "per-tensor") else 1.0
    x = (torch.randn(2, M // 2, K, generator=g) * scale).to(dtype).to(DEV)
    y = mod(x)
    qs = float(mod.quant_scale) if (hasattr(mod, "quant_scale") and act == "per-tensor") else None
    want = need(reference_forward(ref_gemm, x, mod.weight, mod.dequant_scale, mod.bias, act, qs))
    torch.cuda.synchronize()
    assert y.dtype == want.dtype and y.shape == want.shape
    assert torch.equal(y, want), f"max |diff| {(y.float() - want.float()).abs().max().item()}"


BASELINE_SHAPES = [  # Llama-2-7B prefill launches of BASELINE configs[1] (M = 2048 tokens): (K, N, class, granularity)
    (4096, 4096, "W8A8BFP32OFP32Linear", "per-tensor"),                  # q_proj
    (4096, 11008, "W8A8BFP32OFP32Linear", "per-tensor"),                 # gate_proj / up_proj
    (4096, 4096, "W8A8BFP32OFP32LinearWithQuantScale", "per-tensor"),    # o_proj
    (11008, 4096, "W8A8BFP32OFP32LinearWithQuantScale", "per-tensor"),   # down_proj
    (11008, 4096, "W8A8BFP32OFP32LinearWithQuantScale", "per-token"),    # down_proj, BASELINE config 3 granularity
]


@pytest.mark.parametrize("K,N,cls,act", BASELINE_SHAPES)
def test_fused_module_equals_reference_path_at_baseline_sizes(ref_gemm, K, N, cls, act):
    """The FUSED OUTPUT (not only the int32 GEMM) at the sizes the metric is quoted on: one launch of ours against the
    reference's eager prologue + its own cuBLASLt GEMM + eager epilogue, bf16, every element bit-equal."""
    g = torch.Generator(device=DEV).manual_seed(K + N)
    M = 2048
    lin = torch.nn.Linear(K, N, bias=True, device=DEV)
    with torch.no_grad():
        lin.weight.normal_(0.0, 0.02, generator=g)
        lin.bias.normal_(0.0, 1.0, generator=g)
    mod = getattr(NN, cls).from_float(lin, 0.04, save_device=DEV, act_quant=act).to(DEV)
    del lin
    scale = 30.0 if cls == "W8A8BFP32OFP32Linear" else 1.0
    x = (torch.randn(1, M, K, device=DEV, generator=g) * scale).to(torch.bfloat16)
    x[0, 5, :] = 0.25  # a flat row: every element ties at the same code
    y = mod(x)
    qs = float(mod.quant_scale) if (hasattr(mod, "quant_scale") and act == "per-tensor") else None
    want = need(reference_forward(ref_gemm, x, mod.weight, mod.dequant_scale, mod.bias, act, qs))
    torch.cuda.synchronize()
    assert y.dtype == want.dtype and y.shape == want.shape == (1, M, N)
    assert torch.equal(y, want), f"max |diff| {(y.float() - want.float()).abs().max().item()}"
    # and the codes the GEMM consumed are not degenerate: the per-tensor cases saturate some and use the whole range
    assert int(mod.weight.abs().max()) == 127


def test_fused_qkv_module_equals_reference_path_at_baseline_size(ref_gemm):
    """W8A8BFP32OFP32QKVLinear.forward (linear.py:172-208) at 2048 x 12288 x 4096: one int32 GEMM over the packed
    weight, every third dequantised with its own scale, concatenated — against our single launch with a column-scale vector."""
    g = torch.Generator(device=DEV).manual_seed(5)
    M, K, H = 2048, 4096, 4096
    lin = torch.nn.Linear(K, 3 * H, bias=True, device=DEV)
    with torch.no_grad():
        lin.weight.normal_(0.0, 0.02, generator=g)
        lin.weight[H:2 * H] *= 3.0  # the three blocks get visibly different weight scales
        lin.bias.normal_(0.0, 1.0, generator=g)
    mod = NN.W8A8BFP32OFP32QKVLinear.from_float(lin, 0.04, [H, H, H], save_device=DEV, act_quant="per-tensor").to(DEV)
    del lin
    x = (torch.randn(M, K, device=DEV, generator=g) * 30.0).to(torch.bfloat16)
    y = mod(x)
    q = x.round().clamp(-128, 127).to(torch.int8)
    acc = need(ref_o32(ref_gemm, q, mod.weight))
    scales = [float(mod.q_dequant_scale), float(mod.k_dequant_scale), float(mod.v_dequant_scale)]
    assert len({round(s, 12) for s in scales}) == 3
    parts = [s * a for s, a in zip(scales, acc.split([H, H, H], dim=-1))]
    want = (torch.cat(parts, dim=-1) + mod.bias).to(torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(y, want), f"max |diff| {(y.float() - want.float()).abs().max().item()}"
