"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors.

Bars: bit-exact for int8 / int32 / index work and for the fused INT8 outputs (the epilogue
reproduces the reference's fp32 operations one rounding at a time, so even the 16-bit outputs are
required to be identical, which is stricter than the 1-ulp bound BASELINE.md states); FP8 outputs
within the stated accumulation tolerance.  Nothing here reads /root/reference.
"""
import numpy as np
import pytest
import torch

from oracle import w8a8_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from autosmoothquant_b200 import _lib as L
    from autosmoothquant_b200._CUDA import I8CUGEMM
    from autosmoothquant_b200.layers.nn import linear as NN

DEV = torch.device("cuda:0")
TORCH_DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    assert L.load().asq_device_supported() == 1, "GPU is not sm_100 (B200)"


@pytest.fixture
def exact_div():
    prev = L.set_div_mode(L.DIV_EXACT)
    yield
    L.set_div_mode(prev)


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        x = x.to(dtype)
    return x.to(DEV)


def make_x(rng, M, K, dtype, scale=1.0):
    x = rng.standard_normal((M, K)).astype(np.float32) * scale
    x[:, rng.integers(0, K, size=max(1, K // 256))] *= 30.0
    if M > 1:
        x[1] = 0.0  # all-zero row (per-token: scale 0 -> 0/0 -> quantised 0)
    if M > 2:
        x[2, :4] = [1e4, -1e4, 0.5, -2.5]  # saturation + ties
    return O.round_to(x, dtype)


# ----------------------------------------------------------------------------- golden fixtures
def build_int8_module(c):
    cls = getattr(NN, c["cls"])
    N, K = c["weight"].shape
    mod = cls(K, N, "bias" in c, c["act_quant"])
    mod.weight = torch.from_numpy(c["weight"])
    mod.dequant_scale = torch.tensor(c["dequant_scale"], dtype=torch.float32)
    if "quant_scale" in c:
        mod.quant_scale = torch.tensor(c["quant_scale"], dtype=torch.float32)
    if "bias" in c:
        mod.bias = torch.from_numpy(c["bias"])
    return mod.to(DEV)


def test_golden_int8_linear(golden, exact_div):
    cases = [c for c in golden if c["kind"] == "int8_linear"]
    assert len(cases) >= 90
    for c in cases:
        mod = build_int8_module(c)
        y = mod(t(c["x"], TORCH_DT[c["dtype"]]))
        assert y.dtype == TORCH_DT[c["dtype"]] and tuple(y.shape) == c["y"].shape
        np.testing.assert_array_equal(y.float().cpu().numpy(), c["y"], err_msg=f"{c['id']} {c['cls']} {c['act_quant']} {c['dtype']}")


def test_golden_qkv_linear(golden, exact_div):
    for c in (c for c in golden if c["kind"] == "int8_qkv"):
        N, K = c["weight"].shape
        mod = NN.W8A8BFP32OFP32QKVLinear(c["qkv_size"], K, N, "bias" in c, c["act_quant"])
        mod.weight = torch.from_numpy(c["weight"])
        for name, key in (("q_dequant_scale", "q_scale"), ("k_dequant_scale", "k_scale"), ("v_dequant_scale", "v_scale")):
            setattr(mod, name, torch.tensor(c[key], dtype=torch.float32))
        if "bias" in c:
            mod.bias = torch.from_numpy(c["bias"])
        mod = mod.to(DEV)
        y = mod(t(c["x"], TORCH_DT[c["dtype"]]))
        np.testing.assert_array_equal(y.float().cpu().numpy(), c["y"], err_msg=c["id"])


def test_golden_fp8_quantisers(golden, exact_div):
    for c in golden:
        if c["kind"] == "fp8_per_token":
            q, s = L.quantize_act(t(c["x"], TORCH_DT[c["dtype"]]), L.ACT_PER_TOKEN, fp8=True)
            np.testing.assert_array_equal(s.cpu().numpy(), c["scale"], err_msg=c["id"])
        elif c["kind"] == "fp8_static":
            q, _ = L.quantize_act(t(c["x"], TORCH_DT[c["dtype"]]), L.ACT_SCALE, c["in_scale"], fp8=True)
        else:
            continue
        got, want = q.view(torch.uint8).cpu().numpy(), c["q"]
        nan = (want & 0x7F) == 0x7F
        np.testing.assert_array_equal((got & 0x7F) == 0x7F, nan, err_msg=c["id"])
        np.testing.assert_array_equal(got[~nan], want[~nan], err_msg=c["id"])


def test_golden_fp8_linear(golden, exact_div):
    """Tensor-core fp32 accumulation vs the reference's dequantise + fp32 GEMM: both approximate the fp64
    value; tolerance = K * 2^-24 * sum|terms| (bounded here by 2e-5 of the output scale)."""
    for c in (c for c in golden if c["kind"] == "fp8_linear" and c["act"] != "per-tensor" and "out_scale" not in c):
        N, K = c["w"].shape
        if c["act"] == "per-token":
            mod = NN.FP8LinearDynamic(K, N, "per-token", "bias" in c)
        else:
            mod = NN.FP8LinearStatic(K, N, "bias" in c)
            mod.input_scale = torch.tensor(c["in_scale"], dtype=torch.float32)
            mod.output_scale = torch.tensor(0.0)
        mod.weight = torch.from_numpy(c["w"]).view(torch.float8_e4m3fn)
        mod.weight_scale = torch.tensor(c["w_scale"], dtype=torch.float32)
        if "bias" in c:
            mod.bias = torch.from_numpy(c["bias"])
        mod = mod.to(DEV)
        y = mod(t(c["x"])).cpu().numpy()
        np.testing.assert_allclose(y, c["y"], rtol=0, atol=2e-5 * np.abs(c["y"]).max(), err_msg=c["id"])


# ----------------------------------------------------------------------------- int32 GEMM
@pytest.mark.parametrize("M,N,K", [(1, 8, 16), (128, 256, 128), (127, 255, 112), (129, 257, 272), (300, 64, 48),
                                   (5, 130, 1040), (512, 768, 3072), (1000, 1000, 1008), (16, 4096, 4096)])
def test_i8gemm_o32_exact(M, N, K):
    rng = np.random.default_rng(M * 31 + N * 7 + K)
    a = rng.integers(-128, 128, size=(M, K), dtype=np.int8)
    w = rng.integers(-128, 128, size=(N, K), dtype=np.int8)
    a[0, :] = -128
    w[0, :] = -128  # extreme accumulator K * 16384
    out = torch.full((M, N), -7, dtype=torch.int32, device=DEV)
    I8CUGEMM().linear_a8_w8_o32_(t(a), t(w), out)
    np.testing.assert_array_equal(out.cpu().numpy(), O.int8_gemm_i32(a, w))


def test_i8gemm_epi_variants():
    """o8 / bias / relu epilogues (I8CUGEMM.linear_a8_w8_o8*, csrc/kernels/linear.cu variants)."""
    rng = np.random.default_rng(5)
    M, N, K = 70, 96, 160
    a = rng.integers(-128, 128, size=(M, K), dtype=np.int8)
    w = rng.integers(-128, 128, size=(N, K), dtype=np.int8)
    bias8 = rng.integers(-128, 128, size=(N,), dtype=np.int8)
    acc = O.int8_gemm_i32(a, w).astype(np.float32)
    alpha, beta = np.float32(0.0007), np.float32(0.35)
    g = I8CUGEMM()
    out = torch.empty((M, N), dtype=torch.int8, device=DEV)
    g.linear_a8_w8_o8(t(a), t(w), out, float(alpha))
    np.testing.assert_array_equal(out.cpu().numpy(), O.sat_i8(np.rint(alpha * acc)))
    got = g.linear_a8_w8_b8_o8_(t(a), t(w), t(bias8), float(alpha), float(beta))
    want = O.sat_i8(np.rint(alpha * acc + beta * bias8.astype(np.float32)))
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # beta != 0 on the in-place variant: the previous contents of `out` are the addend (bindings.cpp:104-121)
    c0 = rng.integers(-128, 128, size=(M, N), dtype=np.int8)
    out2 = t(c0).clone()
    g.linear_a8_w8_o8_(t(a), t(w), out2, float(alpha), float(beta))
    np.testing.assert_array_equal(out2.cpu().numpy(), O.sat_i8(np.rint(alpha * acc + beta * c0.astype(np.float32))))
    outf = torch.empty((M, N), dtype=torch.float32, device=DEV)
    biasf = rng.standard_normal(N).astype(np.float32)
    L.i8gemm_epi(t(a), t(w), outf, float(alpha), float(beta), bias=t(biasf), relu=True)
    np.testing.assert_array_equal(outf.cpu().numpy(), np.maximum(alpha * acc + beta * biasf, 0))


# ----------------------------------------------------------------------------- prologue tap
@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("mode,name", [(0, "round"), (1, "scale"), (2, "per-token")])
@pytest.mark.parametrize("div", ["exact", "reciprocal"])
def test_quantize_act_int8_vs_oracle(dtype, mode, name, div):
    rng = np.random.default_rng(11)
    for (M, K) in [(3, 16), (37, 272), (130, 4096), (9, 11008)]:
        x = make_x(rng, M, K, dtype, 40.0 if name == "round" else 1.0)
        q, s = L.quantize_act(t(x, TORCH_DT[dtype]), mode, 0.0473, div_mode=L.DIV_EXACT if div == "exact" else L.DIV_RECIPROCAL)
        qw, sw = O.quantize_act_int8(x, dtype, name, 0.0473, div_mode=div)
        np.testing.assert_array_equal(q.cpu().numpy(), qw, err_msg=f"{M}x{K}")
        if sw is not None:
            np.testing.assert_array_equal(s.cpu().numpy(), sw)


@pytest.mark.parametrize("dtype", ["bf16", "f32"])
def test_quantize_act_fp8_vs_oracle(dtype, exact_div):
    rng = np.random.default_rng(12)
    x = make_x(rng, 33, 528, dtype)
    x[1] = rng.standard_normal(528).astype(np.float32) * 1e-4  # tiny row -> subnormal e4m3 codes after scaling? no: per-token rescales
    x = O.round_to(x, dtype)
    for mode, name in ((L.ACT_PER_TOKEN, "per-token"), (L.ACT_SCALE, "scale")):
        q, s = L.quantize_act(t(x, TORCH_DT[dtype]), mode, 3.7, fp8=True)
        qw, sw = O.quantize_act_fp8(x, dtype, name, 3.7)
        np.testing.assert_array_equal(q.view(torch.uint8).cpu().numpy(), qw, err_msg=name)
        if sw is not None:
            np.testing.assert_array_equal(s.cpu().numpy(), sw)


# ----------------------------------------------------------------------------- fused linear
@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("act,qs", [("per-tensor", None), ("per-tensor", 0.0473), ("per-token", None)])
@pytest.mark.parametrize("bias", [False, True])
def test_fused_linear_vs_oracle(dtype, act, qs, bias, exact_div):
    rng = np.random.default_rng(21)
    for (M, N, K) in [(1, 16, 16), (200, 1000, 528), (300, 264, 1040), (129, 4096, 256)]:
        x = make_x(rng, M, K, dtype, 40.0 if (act == "per-tensor" and qs is None) else 1.0)
        w = rng.integers(-127, 128, size=(N, K), dtype=np.int8)
        b = rng.standard_normal(N).astype(np.float32) if bias else None
        mode = L.ACT_PER_TOKEN if act == "per-token" else (L.ACT_ROUND if qs is None else L.ACT_SCALE)
        y = L.w8a8_linear(t(x, TORCH_DT[dtype]), t(w), t(b) if bias else None, mode, qs or 1.0, 0.00321)
        want = O.w8a8_linear(x, dtype, w, 0.00321, act_quant=act, bias=b, quant_scale=qs)
        np.testing.assert_array_equal(y.float().cpu().numpy(), want, err_msg=f"{M}x{N}x{K}")


def test_reciprocal_mode_matches_torch_cuda_eager():
    """Default mode = the reference running on its supported device: torch's CUDA kernels multiply by the
    reciprocal of scalar divisors.  Restate linear.py:278-302 / :83-106 with eager CUDA ops and compare."""
    gen = torch.Generator(device="cpu").manual_seed(3)
    M, N, K = 257, 520, 1040
    for dtype in (torch.bfloat16, torch.float16, torch.float32):
        x = (torch.randn(M, K, generator=gen)).to(dtype).to(DEV)
        w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=gen).to(DEV)
        bias = torch.randn(N, generator=gen).to(DEV)
        qs, ds = 0.0473, 0.00321
        # per-tensor with quant scale
        q = (x / qs).round().clamp(-128, 127).to(torch.int8)
        acc = (q.double() @ w.double().t()).to(torch.int32)
        want = (ds * acc + bias).to(dtype)
        got = L.w8a8_linear(x, w, bias, L.ACT_SCALE, qs, ds, div_mode=L.DIV_RECIPROCAL)
        assert torch.equal(got, want), dtype
        # per-token
        s = x.abs().max(dim=-1, keepdim=True)[0].div(127.0).to(torch.float32)
        q = (x / s).round().clamp(-128, 127).to(torch.int8)
        acc = (q.double() @ w.double().t()).to(torch.int32)
        want = ((ds * s) * acc + bias).to(dtype)
        got = L.w8a8_linear(x, w, bias, L.ACT_PER_TOKEN, 1.0, ds, div_mode=L.DIV_RECIPROCAL)
        assert torch.equal(got, want), dtype


def test_row_scale_given_equals_per_token(exact_div):
    rng = np.random.default_rng(8)
    x = make_x(rng, 150, 528, "bf16")
    w = rng.integers(-127, 128, size=(264, 528), dtype=np.int8)
    xs = t(x, torch.bfloat16)
    rs = torch.empty(150, dtype=torch.float32, device=DEV)
    y1 = L.w8a8_linear(xs, t(w), None, L.ACT_PER_TOKEN, 1.0, 0.01, row_scale_out=rs)
    y2 = L.w8a8_linear(xs, t(w), None, L.ACT_ROW_SCALE_GIVEN, 1.0, 0.01, row_scale_out=rs.clone())
    assert torch.equal(y1, y2)


def test_module_shapes_empty_and_3d(exact_div):
    lin = torch.nn.Linear(64, 48)
    mod = NN.W8A8BFP32OFP32LinearWithQuantScale.from_float(lin, 0.05, act_quant="per-token").to(DEV)
    y = mod(torch.randn(2, 5, 64, device=DEV, dtype=torch.bfloat16))
    assert y.shape == (2, 5, 48) and y.dtype == torch.bfloat16
    y0 = mod(torch.empty(0, 64, device=DEV, dtype=torch.bfloat16))  # empty MoE expert
    assert y0.shape == (0, 48)
    with pytest.raises(RuntimeError):
        mod(torch.randn(3, 64))  # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        L.i8gemm_o32(torch.zeros(4, 24, dtype=torch.int8, device=DEV), torch.zeros(4, 24, dtype=torch.int8, device=DEV),
                     torch.zeros(4, 4, dtype=torch.int32, device=DEV))  # K % 16 != 0


def test_workspace_reuse_across_shapes(exact_div):
    """Regression: the phase counters live at a fixed workspace offset, so alternating shapes on one
    stream must stay exact (stale scratch data must never be read as a 'panel ready' count)."""
    rng = np.random.default_rng(4)
    shapes = [(300, 256, 768), (2048, 512, 1024), (64, 128, 4096), (1000, 384, 272), (300, 256, 768)]
    for (M, N, K) in shapes:
        x = make_x(rng, M, K, "bf16")
        w = rng.integers(-127, 128, size=(N, K), dtype=np.int8)
        for _ in range(2):
            y = L.w8a8_linear(t(x, torch.bfloat16), t(w), None, L.ACT_PER_TOKEN, 1.0, 0.002)
        want = O.w8a8_linear(x, "bf16", w, 0.002, act_quant="per-token")
        np.testing.assert_array_equal(y.float().cpu().numpy(), want, err_msg=f"{M}x{N}x{K}")


# ----------------------------------------------------------------------------- BASELINE sizes: properties
@pytest.mark.parametrize("M,N,K", [(2048, 4096, 4096), (2048, 11008, 4096), (2048, 4096, 11008)])
def test_full_size_properties(M, N, K):
    """Llama-2-7B shapes, checked through size-independent properties (the CPU oracle would take minutes):
    (1) checksum of checksums: sum_n C[m,n] == a[m,:] . (sum_n w[n,:]) in exact integer arithmetic;
    (2) row permutation equivariance; (3) composition: fused output == reference epilogue applied to the
    int32 tap of the prologue tap (all three taps are individually oracle-checked at small sizes)."""
    gen = torch.Generator(device="cpu").manual_seed(M + N + K)
    x = (torch.randn(M, K, generator=gen)).to(torch.bfloat16).to(DEV)
    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=gen).to(DEV)
    bias = torch.randn(N, generator=gen).to(DEV)
    q, s = L.quantize_act(x, L.ACT_PER_TOKEN)
    acc = torch.empty((M, N), dtype=torch.int32, device=DEV)
    L.i8gemm_o32(q, w, acc)
    colsum = w.to(torch.int64).sum(dim=0)  # [K]
    want_rowsum = (q.to(torch.int64) * colsum).sum(dim=1)
    assert torch.equal(acc.to(torch.int64).sum(dim=1), want_rowsum)
    perm = torch.randperm(M, generator=gen).to(DEV)
    acc_p = torch.empty_like(acc)
    L.i8gemm_o32(q[perm].contiguous(), w, acc_p)
    assert torch.equal(acc_p, acc[perm])
    ds = 0.00321
    y = L.w8a8_linear(x, w, bias, L.ACT_PER_TOKEN, 1.0, ds)
    want = ((ds * s.view(-1, 1)) * acc + bias).to(torch.bfloat16)
    assert torch.equal(y, want)
    # per-tensor static: same composition
    qs = 0.0473
    q2, _ = L.quantize_act(x, L.ACT_SCALE, qs)
    L.i8gemm_o32(q2, w, acc)
    y2 = L.w8a8_linear(x, w, None, L.ACT_SCALE, qs, ds)
    assert torch.equal(y2, (ds * acc).to(torch.bfloat16))


# ----------------------------------------------------------------------------- model-level plumbing
def test_fused_projections_equal_separate_projections():
    """The benchmark stack fuses q|k|v and gate|up through W8A8BFP32OFP32QKVLinear; per-block dequant scales
    make that bit-identical to one launch per projection."""
    from autosmoothquant_b200 import harness

    ids = torch.randint(0, harness.TINY.vocab, (2, 64), generator=torch.Generator().manual_seed(0)).to(DEV)
    for qc in ({}, {"qkv": "per-token", "out": "per-token", "fc1": "per-token", "fc2": "per-token"}):
        a = harness.QuantDecoder(harness.TINY, qc, device=DEV, seed=3, fuse_projections=False)
        b = harness.QuantDecoder(harness.TINY, qc, device=DEV, seed=3, fuse_projections=True)
        ya, yb = a(ids, last_token_only=False), b(ids, last_token_only=False)
        assert torch.isfinite(ya).all()
        assert torch.equal(ya, yb)


def test_decoder_layer_matches_oracle_linears(exact_div):
    """One decoder layer's quantized projections, module by module, against the oracle on the activations the
    layer actually produces (per-tensor config: round-only qkv/fc1, quant-scale out/fc2)."""
    from autosmoothquant_b200 import harness

    model = harness.QuantDecoder(harness.TINY, {}, device=DEV, seed=5)
    layer = model.layers[0]
    captured = {}
    hooks = [m.register_forward_hook(lambda mod, inp, out, name=n: captured.__setitem__(name, (inp[0], out)))
             for n, m in layer.named_children()]
    ids = torch.randint(0, harness.TINY.vocab, (1, 48), generator=torch.Generator().manual_seed(1)).to(DEV)
    model(ids)
    for h in hooks:
        h.remove()
    assert set(captured) == {"q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"}
    for name, (x, y) in captured.items():
        mod = getattr(layer, name)
        want = O.w8a8_linear(x.float().cpu().numpy(), "bf16", mod.weight.cpu().numpy(), float(mod.dequant_scale),
                             act_quant="per-tensor",
                             quant_scale=float(mod.quant_scale) if hasattr(mod, "quant_scale") else None)
        np.testing.assert_array_equal(y.float().cpu().numpy(), want, err_msg=name)


# ----------------------------------------------------------------------------- producer-side fusions (asq_glue.cu)
def _frac_bad(got, want):
    return float(np.mean(got != want))


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("H", [256, 4096, 5120, 8192])
def test_add_rmsnorm_quant(dtype, H):
    """fp32 row reduction: the summation order differs from the oracle's fp64 mean, so a last-place flip of
    rstd may move a handful of outputs by one ulp / one int8 step; everything else must be identical."""
    rng = np.random.default_rng(H)
    M = 67
    x = O.round_to(rng.standard_normal((M, H)).astype(np.float32) * 2, dtype)
    d = O.round_to(rng.standard_normal((M, H)).astype(np.float32), dtype)
    w = O.round_to((1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32) / 0.035, dtype)  # folded 1/input_scale
    for delta in (None, d):
        xs, h, q = L.add_rmsnorm_quant(t(x, TORCH_DT[dtype]), None if delta is None else t(delta, TORCH_DT[dtype]),
                                       t(w, TORCH_DT[dtype]), 1e-5, want_h=True, want_q=True)
        xw, hw, qw = O.add_rmsnorm_quant(x, delta, w, 1e-5, dtype)
        np.testing.assert_array_equal(xs.float().cpu().numpy(), xw)
        hg = h.float().cpu().numpy()
        assert _frac_bad(hg, hw) < 5e-3
        np.testing.assert_allclose(hg, hw, rtol=2 ** -6 if dtype == "bf16" else 2 ** -9, atol=0)  # <= 2 ulp (two roundings)
        qg = q.cpu().numpy().astype(np.int32)
        assert np.abs(qg - qw.astype(np.int32)).max() <= 1 and _frac_bad(qg, qw) < 5e-3
        # the int8 output is exactly the rounding of the kernel's own h (what Linear.forward would compute)
        np.testing.assert_array_equal(qg, O.sat_i8(np.rint(hg)))


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_silu_mul_quant(dtype, exact_div):
    rng = np.random.default_rng(3)
    M, I = 45, 1376
    gu = O.round_to(rng.standard_normal((M, 2 * I)).astype(np.float32) * 2, dtype)
    q, a = L.silu_mul_quant(t(gu, TORCH_DT[dtype]), 0.0631, want_q=True, want_a=True)
    aw, qw = O.silu_mul_quant(gu[:, :I], gu[:, I:], 0.0631, dtype)
    ag = a.float().cpu().numpy()
    assert _frac_bad(ag, aw) < 2e-3  # expf vs fp64 exp: rare one-ulp differences before rounding to T
    np.testing.assert_allclose(ag, aw, rtol=2 ** -6 if dtype == "bf16" else 2 ** -9, atol=1e-30)
    qg = q.cpu().numpy().astype(np.int32)
    assert np.abs(qg - qw.astype(np.int32)).max() <= 1 and _frac_bad(qg, qw) < 2e-3
    # int8 output == the module path applied to the kernel's own activation
    np.testing.assert_array_equal(qg, O.quantize_act_int8(ag, dtype, "scale", 0.0631)[0])


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
def test_rope_inplace_bit_exact(dtype):
    rng = np.random.default_rng(9)
    B, S, nq, nk, hd = 2, 37, 6, 2, 64
    row = (nq + 2 * nk) * hd
    qkv = O.round_to(rng.standard_normal((B * S, row)).astype(np.float32), dtype)
    ang = np.outer(np.arange(S), 1.0 / (10000.0 ** (np.arange(0, hd, 2) / hd))).astype(np.float32)
    emb = np.concatenate([ang, ang], axis=-1)
    cos, sin = O.round_to(np.cos(emb), dtype), O.round_to(np.sin(emb), dtype)
    buf = t(qkv, TORCH_DT[dtype])
    L.rope_inplace(buf, t(cos, TORCH_DT[dtype]), t(sin, TORCH_DT[dtype]), S, nq + nk, hd)
    got = buf.float().cpu().numpy()
    want = qkv.copy()
    heads = qkv[:, :(nq + nk) * hd].reshape(B, S, nq + nk, hd).transpose(0, 2, 1, 3)  # [B, heads, S, hd]
    rot = O.rope_rotate_half(heads, cos, sin, dtype).transpose(0, 2, 1, 3).reshape(B * S, -1)
    want[:, :(nq + nk) * hd] = rot
    np.testing.assert_array_equal(got, want)  # v block untouched, q/k rotated exactly


def test_linear_q8_equals_module_path(exact_div):
    """Pre-quantised entry point == fused entry point when fed the same int8 activations."""
    rng = np.random.default_rng(6)
    M, N, K = 300, 520, 1040
    x = make_x(rng, M, K, "bf16", 40.0)
    w = rng.integers(-127, 128, size=(N, K), dtype=np.int8)
    b = rng.standard_normal(N).astype(np.float32)
    y1 = L.w8a8_linear(t(x, torch.bfloat16), t(w), t(b), L.ACT_ROUND, 1.0, 0.004)
    q, _ = L.quantize_act(t(x, torch.bfloat16), L.ACT_ROUND)
    y2 = L.w8a8_linear_q8(q, t(w), t(b), 0.004)
    assert torch.equal(y1, y2)
    y3 = L.w8a8_linear(t(x, torch.bfloat16), t(w), t(b), L.ACT_PER_TOKEN, 1.0, 0.004)
    q, s = L.quantize_act(t(x, torch.bfloat16), L.ACT_PER_TOKEN)
    y4 = L.w8a8_linear_q8(q, t(w), t(b), 0.004, row_scale=s)
    assert torch.equal(y3, y4)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("M,I,K,per_token,bias", [(300, 96, 256, False, True), (512, 1376, 1024, False, False),
                                                  (77, 160, 512, True, True), (2048, 11008, 4096, False, False)])
@pytest.mark.parametrize("div", ["exact", "reciprocal"])
def test_gateup_swiglu_epilogue_equals_two_launch_path(dtype, M, I, K, per_token, bias, div):
    """SiLU(gate)*up (+ down_proj's per-tensor quantisation) in the gate|up GEMM epilogue must emit exactly the
    bytes of the two-launch path it replaces (fused gate|up GEMM -> silu_mul_quant), which in turn is tied to
    the oracle by test_silu_mul_quant / test_fused_linear_vs_oracle.  Covers M / N tile tails (I = 96 -> one
    192-wide tile, I = 160 -> 256 + 64), per-token input scales and bias."""
    if M * I * K > 2 ** 33 and (dtype == "f16" or div == "exact"):
        pytest.skip("full-size case runs once")
    g = torch.Generator().manual_seed(I + M)
    td = TORCH_DT[dtype]
    xq = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).to(DEV)
    wg = torch.randint(-127, 128, (I, K), dtype=torch.int8, generator=g).to(DEV)
    wu = torch.randint(-127, 128, (I, K), dtype=torch.int8, generator=g).to(DEV)
    sg, su = 2.9e-6 * (4096 / K) ** 0.5, 4.3e-6 * (4096 / K) ** 0.5  # gate / up keep their own dequant scale
    cs = torch.cat([torch.full((I,), sg), torch.full((I,), su)]).to(DEV)
    b = (torch.randn(2 * I, generator=g) * 0.5).to(DEV) if bias else None
    rs = (torch.rand(M, generator=g) * 2 + 0.5).to(DEV) if per_token else None
    qs = 0.0431
    mode = L.DIV_EXACT if div == "exact" else L.DIV_RECIPROCAL
    gu = L.w8a8_linear_q8(xq, torch.cat([wg, wu]).contiguous(), b, 1.0, col_scale=cs, row_scale=rs, out_dtype=td)
    q_want, a_want = L.silu_mul_quant(gu, qs, want_q=True, want_a=True, div_mode=mode)
    w_il = L.interleave_gate_up(wg, wu)
    cs_il = L.interleave_gate_up(cs[:I], cs[I:])
    b_il = L.interleave_gate_up(b[:I], b[I:]) if bias else None
    q_got = L.w8a8_gateup_swiglu(xq, w_il, b_il, 1.0, col_scale_il=cs_il, row_scale=rs, out_quant_scale=qs,
                                 mid_dtype=td, div_mode=mode)
    a_got = L.w8a8_gateup_swiglu(xq, w_il, b_il, 1.0, col_scale_il=cs_il, row_scale=rs, out_quant_scale=None, mid_dtype=td)
    torch.cuda.synchronize()
    assert q_got.dtype == torch.int8 and q_got.shape == (M, I) and a_got.dtype == td
    assert float(a_want.float().abs().max()) > 0.5 and int(q_want.abs().max()) > 20  # the data exercise the range
    assert torch.equal(a_got, a_want), f"product differs in {(a_got != a_want).sum().item()} elements"
    assert torch.equal(q_got, q_want), f"int8 differs in {(q_got != q_want).sum().item()} elements"
    # the two scalar dequant scales instead of the per-column vector
    q2 = L.w8a8_gateup_swiglu(xq, w_il, b_il, sg, up_dequant_scale=su, row_scale=rs, out_quant_scale=qs, mid_dtype=td,
                              div_mode=mode)
    assert torch.equal(q2, q_want)
    q3 = L.w8a8_gateup_swiglu(xq, w_il, b_il, sg, row_scale=rs, out_quant_scale=qs, mid_dtype=td, div_mode=mode)
    gu3 = L.w8a8_linear_q8(xq, torch.cat([wg, wu]).contiguous(), b, sg, row_scale=rs, out_dtype=td)
    assert torch.equal(q3, L.silu_mul_quant(gu3, qs, div_mode=mode)[0])


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("M,S,nq,nk,nv,K,per_token,bias", [(300, 150, 3, 1, 1, 512, False, True), (64, 64, 2, 2, 2, 256, True, False),
                                                           (2048, 2048, 32, 32, 32, 4096, False, False)])
def test_rope_epilogue_equals_gemm_then_rope_kernel(dtype, M, S, nq, nk, nv, K, per_token, bias):
    """RoPE in the q|k|v GEMM epilogue == the GEMM followed by asq_rope_inplace (itself bit-exact against the
    oracle's HF rotate-half, test_rope_inplace_bit_exact); v columns must be untouched."""
    if M == 2048 and dtype == "f16":
        pytest.skip("full-size case runs once")
    hd = 128
    g = torch.Generator().manual_seed(M + nq)
    td = TORCH_DT[dtype]
    N = (nq + nk + nv) * hd
    xq = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).to(DEV)
    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g).to(DEV)
    sc = 3e-6 * (4096 / K) ** 0.5
    cs = torch.cat([torch.full((nq * hd,), sc), torch.full((nk * hd,), 1.3 * sc), torch.full((nv * hd,), 0.7 * sc)]).to(DEV)
    b = torch.randn(N, generator=g).to(DEV) if bias else None
    rs = (torch.rand(M, generator=g) + 0.5).to(DEV) if per_token else None
    ang = torch.outer(torch.arange(S, dtype=torch.float32), 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd)))
    emb = torch.cat([ang, ang], dim=-1)
    cos, sin = emb.cos().to(td).to(DEV), emb.sin().to(td).to(DEV)
    want = L.w8a8_linear_q8(xq, w, b, 1.0, col_scale=cs, row_scale=rs, out_dtype=td)
    plain = want.clone()
    L.rope_inplace(want, cos, sin, S, nq + nk, hd)
    got = L.w8a8_linear_q8(xq, w, b, 1.0, col_scale=cs, row_scale=rs, out_dtype=td, rope=(L.rope_tables_blocked(cos), L.rope_tables_blocked(sin), S, (nq + nk) * hd))
    torch.cuda.synchronize()
    assert not torch.equal(want, plain)  # the rotation did something
    assert torch.equal(got[:, (nq + nk) * hd:], plain[:, (nq + nk) * hd:])
    assert torch.equal(got, want), f"{(got != want).sum().item()} elements differ"
    got2 = L.w8a8_linear_q8(xq, w, b, 1.0, col_scale=cs, row_scale=rs, out_dtype=td,
                            rope=(L.rope_tables_blocked(cos), L.rope_tables_blocked(sin), S, (nq + nk) * hd, True))
    assert torch.equal(got2, want)  # HF tables repeat their first half: the halves_equal fast path is identical


def test_rope_epilogue_rejects_unsupported_head_dim():
    xq = torch.zeros((4, 64), dtype=torch.int8, device=DEV)
    w = torch.zeros((128, 64), dtype=torch.int8, device=DEV)
    t64 = L.rope_tables_blocked(torch.zeros((8, 64), dtype=torch.bfloat16, device=DEV))
    with pytest.raises(RuntimeError, match="head_dim"):
        L.w8a8_linear_q8(xq, w, None, 1.0, rope=(t64, t64, 8, 128))


def test_gateup_swiglu_rejects_bad_arguments():
    xq = torch.zeros((4, 64), dtype=torch.int8, device=DEV)
    w = torch.zeros((96, 64), dtype=torch.int8, device=DEV)  # N = 96 is not a multiple of 64
    with pytest.raises(RuntimeError, match="multiple of 64"):
        L.w8a8_gateup_swiglu(xq, w, None, 1.0, out_quant_scale=0.1)
    w = torch.zeros((128, 64), dtype=torch.int8, device=DEV)
    with pytest.raises(RuntimeError, match="positive"):
        L.w8a8_gateup_swiglu(xq, w, None, 1.0, out_quant_scale=0.0)
    assert L.w8a8_gateup_swiglu(xq[:0], w, None, 1.0, out_quant_scale=0.1).shape == (0, 64)


def test_glue_stack_swiglu_epilogue_is_bit_identical():
    """The decoder forward with the SwiGLU epilogue == the same forward with the separate SiLU kernel."""
    from autosmoothquant_b200 import harness

    ids = torch.randint(0, harness.TINY.vocab, (2, 96), generator=torch.Generator().manual_seed(1)).to(DEV)
    a = harness.QuantDecoder(harness.TINY, {}, device=DEV, seed=5, fuse_projections=True, glue=True, swiglu_epilogue=False)
    b = harness.QuantDecoder(harness.TINY, {}, device=DEV, seed=5, fuse_projections=True, glue=True, swiglu_epilogue=True)
    assert hasattr(b.layers[0], "gate_up_il") and not hasattr(a.layers[0], "gate_up_il")
    assert torch.equal(a(ids, last_token_only=False), b(ids, last_token_only=False))


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("M,N,K", [(300, 520, 256), (2048, 4096, 4096), (64, 72, 64)])
def test_residual_epilogue_equals_linear_then_add(dtype, M, N, K, exact_div):
    """y = T(residual + T(linear(x))) from one launch == the linear launch followed by the eager add (what the
    decoder block's `hidden = residual + o_proj(...)` computes), for the fused (bf16 in) and the int8-in entry."""
    if M == 2048 and dtype == "f16":
        pytest.skip("full-size case runs once")
    td = TORCH_DT[dtype]
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).to(td).to(DEV)
    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    res = (torch.randn(M, N, generator=g) * 3).to(td).to(DEV)
    for mode, qs in ((L.ACT_SCALE, 0.031), (L.ACT_PER_TOKEN, 1.0)):
        plain = L.w8a8_linear(x, w, b, mode, qs, 2.1e-4)
        got = L.w8a8_linear(x, w, b, mode, qs, 2.1e-4, residual=res)
        assert torch.equal(got, res + plain)
    xq, _ = L.quantize_act(x, L.ACT_SCALE, 0.031)
    plain = L.w8a8_linear_q8(xq, w, None, 2.1e-4, out_dtype=td)
    got = L.w8a8_linear_q8(xq, w, None, 2.1e-4, out_dtype=td, residual=res)
    assert torch.equal(got, res + plain)
    inplace = res.clone()  # the residual may alias the output (the decoder updates its stream in place)
    rc = L.load().asq_w8a8_linear_q8_res(xq.data_ptr(), None, w.data_ptr(), None, inplace.data_ptr(), inplace.data_ptr(), L._code(td),
                                         M, N, K, 2.1e-4, None, torch.cuda.current_stream().cuda_stream)
    assert rc == 0 and torch.equal(inplace, got)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("M,K,nh,I", [(300, 256, 2, 96), (2048, 4096, 4, 512), (130, 5120, 1, 64), (64, 1008, 1, 32)])
def test_rmsnorm_prologue_equals_norm_kernel_then_int8_entry(dtype, M, K, nh, I):
    """RMSNorm as the prologue of the q|k|v (+RoPE) and gate|up (+SwiGLU) launches == asq_add_rmsnorm_quant followed by
    the int8-in entry points, bit for bit (the prologue reproduces the norm kernel's reduction order).  K = 5120
    exercises the two-pass path of rows longer than one register batch, K = 1008 a ragged last vector group."""
    if M == 2048 and dtype == "f16":
        pytest.skip("full-size case runs once")
    td = TORCH_DT[dtype]
    g = torch.Generator().manual_seed(M + K)
    x = (torch.randn(M, K, generator=g) * 1.7).to(td).to(DEV)
    nw = ((torch.rand(K, generator=g) + 0.5) * 25.0).to(td).to(DEV)  # norm weight with a folded 1/input_scale
    eps = 1e-5
    hd = 128
    N = 3 * nh * hd
    wq = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    cs = (torch.rand(N, generator=g) * 2e-5 + 1e-5).to(DEV)
    S = M
    ang = torch.outer(torch.arange(S, dtype=torch.float32), 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd)))
    emb = torch.cat([ang, ang], dim=-1)
    rope = (L.rope_tables_blocked(emb.cos().to(td).to(DEV)), L.rope_tables_blocked(emb.sin().to(td).to(DEV)), S, 2 * nh * hd, True)
    _, _, q8 = L.add_rmsnorm_quant(x, None, nw, eps)
    assert int(q8.abs().max()) > 60
    for r in (None, rope):
        want = L.w8a8_linear_q8(q8, wq, b, 1.0, col_scale=cs, out_dtype=td, rope=r)
        got = L.w8a8_rmsnorm_linear(x, nw, eps, wq, b, 1.0, col_scale=cs, rope=r)
        assert torch.equal(got, want), f"rope={r is not None}: {(got != want).sum().item()} elements differ"
    wg = torch.randint(-127, 128, (I, K), dtype=torch.int8, generator=g).to(DEV)
    wu = torch.randint(-127, 128, (I, K), dtype=torch.int8, generator=g).to(DEV)
    w_il = L.interleave_gate_up(wg, wu)
    sc = 3e-6 * (4096 / K) ** 0.5
    for qs in (0.0431, None):
        want = L.w8a8_gateup_swiglu(q8, w_il, None, sc, up_dequant_scale=1.4 * sc, out_quant_scale=qs, mid_dtype=td)
        got = L.w8a8_rmsnorm_gateup_swiglu(x, nw, eps, w_il, None, sc, up_dequant_scale=1.4 * sc, out_quant_scale=qs)
        assert torch.equal(got, want)


def test_glue_stack_norm_prologue_is_bit_identical(monkeypatch):
    from autosmoothquant_b200 import harness

    ids = torch.randint(0, harness.TINY.vocab, (2, 96), generator=torch.Generator().manual_seed(4)).to(DEV)
    for qc in ({}, {"out": "per-token", "fc2": "per-token"}):
        model = harness.QuantDecoder(harness.TINY, qc, device=DEV, seed=5, fuse_projections=True, glue=True)
        monkeypatch.setenv("ASQ_OPROJ_SPLIT", "0")  # count the one-launch-per-GEMM form (the default quantises o_proj's input in its own launch)
        monkeypatch.setenv("ASQ_NORM_PROLOGUE", "0")
        want = model(ids, last_token_only=False)
        monkeypatch.setenv("ASQ_NORM_PROLOGUE", "1")
        before = L.launch_count()
        got = model(ids, last_token_only=False)
        assert torch.equal(got, want)
        # four GEMM launches per layer (+ the RoPE kernel: the tiny config's head_dim is 64) + the final norm
        per_layer = 4 + (0 if harness.TINY.head_dim == 128 else 1)
        assert L.launch_count() - before == per_layer * len(model.layers) + 1


def test_glue_stack_residual_epilogue_is_bit_identical(monkeypatch):
    from autosmoothquant_b200 import harness

    ids = torch.randint(0, harness.TINY.vocab, (2, 96), generator=torch.Generator().manual_seed(4)).to(DEV)
    for qc in ({}, {"out": "per-token", "fc2": "per-token"}):
        model = harness.QuantDecoder(harness.TINY, qc, device=DEV, seed=5, fuse_projections=True, glue=True)
        monkeypatch.setenv("ASQ_RESIDUAL_EPILOGUE", "0")
        want = model(ids, last_token_only=False)
        monkeypatch.setenv("ASQ_RESIDUAL_EPILOGUE", "1")
        assert torch.equal(model(ids, last_token_only=False), want)
        assert torch.equal(model(ids, last_token_only=True), want[:, -1:, :])


def test_glue_stack_config3_per_token_out_fc2():
    """BASELINE config 3 granularities (qkv / fc1 per-tensor, out / fc2 per-token): the producer-fused path keeps
    the norm->int8 and SwiGLU-epilogue fusions and must agree with the module path like the all-per-tensor case."""
    from autosmoothquant_b200 import harness

    qc = {"qkv": "per-tensor", "out": "per-token", "fc1": "per-tensor", "fc2": "per-token"}
    ids = torch.randint(0, harness.TINY.vocab, (2, 64), generator=torch.Generator().manual_seed(0)).to(DEV)
    a = harness.QuantDecoder(harness.TINY, qc, device=DEV, seed=3, fuse_projections=True, glue=False)
    b = harness.QuantDecoder(harness.TINY, qc, device=DEV, seed=3, fuse_projections=True, glue=True)
    assert b.glue and not a.glue and b.layers[0].down_proj.act_quant == "per-token"
    ya, yb = a(ids, last_token_only=False), b(ids, last_token_only=False)
    assert torch.isfinite(yb).all() and float((ya - yb).abs().max()) <= 0.05 * float(ya.abs().max())
    c = harness.QuantDecoder(harness.TINY, qc, device=DEV, seed=3, fuse_projections=True, glue=True, swiglu_epilogue=False)
    assert torch.equal(yb, c(ids, last_token_only=False))  # SwiGLU epilogue (16-bit output) == separate SiLU kernel


def test_glue_stack_close_to_module_stack():
    """Whole tiny decoder: producer-fused path vs module path.  Not bit-identical by construction (the fp32
    variance is summed in a different order), so compare logits with a tolerance."""
    from autosmoothquant_b200 import harness

    ids = torch.randint(0, harness.TINY.vocab, (2, 64), generator=torch.Generator().manual_seed(0)).to(DEV)
    a = harness.QuantDecoder(harness.TINY, {}, device=DEV, seed=3, fuse_projections=True, glue=False)
    b = harness.QuantDecoder(harness.TINY, {}, device=DEV, seed=3, fuse_projections=True, glue=True)
    assert b.glue
    for last in (False, True):
        ya, yb = a(ids, last_token_only=last), b(ids, last_token_only=last)
        assert ya.shape == yb.shape and torch.isfinite(yb).all()
        assert float((ya - yb).abs().max()) <= 0.05 * float(ya.abs().max())
        assert float((ya != yb).float().mean()) < 0.5


# ----------------------------------------------------------------------------- FP8: remaining reference branches
def test_golden_fp8_per_tensor_dynamic_and_output_fakequant(golden, exact_div):
    """FP8LinearDynamic's per-tensor branch (whole-tensor absmax reduced in-kernel, linear.py:417-418) and
    FP8LinearStatic with a truthy output_scale (epilogue fake-quantisation, linear.py:562-564) against the
    reference's recorded outputs."""
    n = 0
    for c in (c for c in golden if c["kind"] == "fp8_linear" and (c["act"] == "per-tensor" or "out_scale" in c)):
        N, K = c["w"].shape
        dt = TORCH_DT[c["dtype"]]
        if c["act"] == "per-tensor":
            mod = NN.FP8LinearDynamic(K, N, "per-tensor", "bias" in c)
        else:
            mod = NN.FP8LinearStatic(K, N, "bias" in c)
            mod.input_scale = torch.tensor(c["in_scale"], dtype=torch.float32)
            mod.output_scale = torch.tensor(c["out_scale"], dtype=torch.float32)
        mod.weight = torch.from_numpy(c["w"]).view(torch.float8_e4m3fn)
        mod.weight_scale = torch.tensor(c["w_scale"], dtype=torch.float32)
        if "bias" in c:
            mod.bias = torch.from_numpy(c["bias"])
        mod = mod.to(DEV)
        y = mod(t(c["x"], dt)).float().cpu().numpy()
        scale = np.abs(c["y"]).max()
        if "out_scale" in c:
            # outputs live on the e4m3 grid (times out_scale): a value may land on a neighbouring code
            assert np.mean(y != c["y"]) < 0.02 and np.abs(y - c["y"]).max() <= 0.13 * scale, c["id"]
            grid = O.e4m3_decode(np.arange(256, dtype=np.uint8))
            grid = grid[np.isfinite(grid)] * np.float32(c["out_scale"])
            assert np.isin(y, grid).all()
        else:
            tol = 2e-5 if c["dtype"] == "f32" else 2 ** -7
            np.testing.assert_allclose(y, c["y"], rtol=0, atol=tol * scale, err_msg=c["id"])
        n += 1
    assert n == 4


def test_fp8_per_tensor_dynamic_scale_is_the_oracle_scale(exact_div):
    rng = np.random.default_rng(2)
    M, N, K = 700, 264, 528
    x = make_x(rng, M, K, "bf16")
    w = O.e4m3_encode(np.clip(rng.standard_normal((N, K)).astype(np.float32) * 20, -448, 448))
    rs = torch.empty(M, dtype=torch.float32, device=DEV)
    for _ in range(2):  # second launch re-uses the (restored) reduction words
        y = L.fp8_linear(t(x, torch.bfloat16), t(w).view(torch.float8_e4m3fn), None, L.ACT_PER_TENSOR_DYNAMIC, 1.0, 0.01,
                         row_scale_out=rs)
    q, s = O.quantize_act_fp8(x, "bf16", "per-tensor")
    assert torch.all(rs == float(np.asarray(s).reshape(()))).item()
    want = O.fp8_linear_exact(q, w, np.full(M, float(np.asarray(s).reshape(())), np.float32), 0.01)
    ok = np.isfinite(want)
    got = y.float().cpu().numpy()
    np.testing.assert_allclose(got[ok], want[ok], rtol=2 ** -7, atol=2 ** -7 * np.abs(want[ok]).max())


# ----------------------------------------------------------------------------- BASELINE configs 3-5: per-rank shapes
@pytest.mark.parametrize("name,M,N,K,act", [
    ("13B o_proj per-token", 4096, 5120, 5120, "per-token"),
    ("13B down_proj per-token", 4096, 5120, 13824, "per-token"),
    ("13B qkv per-tensor (batch-32 prefill slice)", 8192, 15360, 5120, "per-tensor"),
    ("Mixtral TP8 w1|w3 per-token", 512, 3584, 4096, "per-token"),
    ("Mixtral TP8 w2 per-token (row shard)", 512, 4096, 1792, "per-token"),
    ("Mixtral expert with 3 tokens", 3, 4096, 1792, "per-token"),
])
def test_config_shapes_int8_composition(name, M, N, K, act):
    """Shapes of BASELINE configs 3 and 4 (per rank): fused output == reference epilogue applied to the exact
    int32 GEMM of the prologue tap (each tap is oracle-checked bit-for-bit at small sizes), plus the integer
    checksum-of-checksums of the GEMM."""
    gen = torch.Generator(device="cpu").manual_seed(M + N + K)
    x = torch.randn(M, K, generator=gen).to(torch.bfloat16).to(DEV)
    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=gen).to(DEV)
    mode = L.ACT_PER_TOKEN if act == "per-token" else L.ACT_ROUND
    if act == "per-tensor":
        x = (x.float() * 40).to(torch.bfloat16)
    q, s = L.quantize_act(x, mode)
    acc = torch.empty((M, N), dtype=torch.int32, device=DEV)
    L.i8gemm_o32(q, w, acc)
    assert torch.equal(acc.to(torch.int64).sum(dim=1), (q.to(torch.int64) * w.to(torch.int64).sum(dim=0)).sum(dim=1)), name
    ds = 0.0021
    y = L.w8a8_linear(x, w, None, mode, 1.0, ds)
    want = ((ds * s.view(-1, 1)) * acc if s is not None else ds * acc).to(torch.bfloat16)
    assert torch.equal(y, want), name


@pytest.mark.parametrize("name,M,N,K", [
    ("70B TP8 q_proj", 2048, 1024, 8192), ("70B TP8 kv_proj", 2048, 128, 8192), ("70B TP8 gate|up", 2048, 7168, 8192),
    ("70B TP8 o_proj (row shard)", 2048, 8192, 1024), ("70B TP8 down_proj (row shard)", 2048, 8192, 3584),
])
def test_config5_shapes_fp8_per_token(name, M, N, K):
    """BASELINE config 5 (Llama-2-70B FP8-e4m3 per-token, TP=8) per-rank shapes: tensor-core result vs an fp64
    evaluation of the same quantised operands (the prologue tap is bit-exact vs the oracle at small sizes)."""
    gen = torch.Generator(device="cpu").manual_seed(N + K)
    x = torch.randn(M, K, generator=gen).to(torch.bfloat16).to(DEV)
    wf = torch.randn(N, K, generator=gen) * 0.02
    ws = float(wf.abs().max() / 448.0)
    w = (wf / ws).clamp(-448, 448).to(torch.float8_e4m3fn).to(DEV)
    y = L.fp8_linear(x, w, None, L.ACT_PER_TOKEN, 1.0, ws)
    q, s = L.quantize_act(x, L.ACT_PER_TOKEN, fp8=True)
    want = (q.double() @ w.double().t()) * (s.double().view(-1, 1) * ws)
    err = (y.double() - want).abs().max().item()
    # bf16 output rounding (2^-9 relative) + fp32 accumulation over K terms
    assert err <= want.abs().max().item() * (2 ** -8 + K * 2 ** -24), name


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("use_bias", [False, True])
def test_fp8_fused_columns_equal_separate_launches(dtype, use_bias):
    """q|k|v (and gate|up) FP8 per-token projections as ONE launch with a per-column weight-scale vector
    (asq_fp8_linear_cs, harness.FP8FusedColumnsLinear) == the three module forwards, bit for bit: every output element
    sees the same codes, the same fp32 accumulation and the same factor fl(w_scale * s[m])."""
    from autosmoothquant_b200.harness import FP8FusedColumnsLinear

    g = torch.Generator().manual_seed(3)
    M, K = 300, 512
    mods = []
    for n, std in ((256, 0.02), (128, 0.3), (136, 1.5)):  # three clearly different weight scales; 136: a ragged N tail
        lin = torch.nn.Linear(K, n, bias=use_bias)
        with torch.no_grad():
            lin.weight.copy_(torch.randn(n, K, generator=g) * std)
            if use_bias:
                lin.bias.copy_(torch.randn(n, generator=g))
        mods.append(NN.FP8LinearDynamic.from_float(lin, act_quant="per-token", reference_compat=False)._apply(lambda t: t.to(DEV)))
    assert len({float(m.weight_scale) for m in mods}) == 3
    fused = FP8FusedColumnsLinear(mods)
    x = torch.randn(2, M // 2, K, generator=g)
    x[0, 3, :] = 0.0  # an all-zero token: scale 0
    x[1, 7, 5] = 3e4
    x = x.to(dtype).to(DEV)
    before = L.launch_count()
    y = fused(x)
    assert L.launch_count() - before == 1
    want = torch.cat([m(x) for m in mods], dim=-1)
    torch.cuda.synchronize()
    assert y.dtype == dtype and y.shape == want.shape == (2, M // 2, 520)
    assert torch.equal(y.view(torch.int32 if dtype == torch.float32 else torch.int16),
                       want.view(torch.int32 if dtype == torch.float32 else torch.int16))  # bit patterns: NaN-safe (zero token)
    # the column vector is validated: static scales and wrong lengths are rejected before any launch
    with pytest.raises(ValueError):
        L.fp8_linear(x.view(-1, K), fused.weight, None, L.ACT_PER_TOKEN, col_scale=fused.col_scale[:-1].contiguous())
    with pytest.raises(ValueError):
        L.fp8_linear(x.view(-1, K), fused.weight, None, L.ACT_SCALE, 0.5, col_scale=fused.col_scale)


# ----------------------------------------------------------------------------- batched INT8 GEMM (csrc/kernels/bmm.cu)
@pytest.mark.parametrize("B,M,N,K", [(4, 256, 192, 128), (3, 512, 512, 64), (5, 100, 72, 48), (1, 300, 64, 32), (32, 256, 256, 128)])
def test_i8bmm_family_vs_exact_integer_matmul(B, M, N, K):
    """bmm_s8t_s8n_{s32t,f32t,s8t} (layers/nn/bmm.py, csrc/kernels/bmm.cu:10-211) against exact integer arithmetic:
    int32 exact; float32 = fl(alpha * f32(acc)); int8 = sat(rint(that)).  M multiples of 256 take the one-launch
    stacked path, the others one launch per batch entry."""
    from autosmoothquant_b200.layers.nn.bmm import BMM_S8T_S8N_F32T, BMM_S8T_S8N_S8T, BMM_S8T_S8N_S32T

    rng = np.random.default_rng(B * 1000 + M)
    a = rng.integers(-128, 128, size=(B, M, K), dtype=np.int8)
    b = rng.integers(-128, 128, size=(B, N, K), dtype=np.int8)
    acc = np.einsum("bmk,bnk->bmn", a.astype(np.int64), b.astype(np.int64))
    assert np.abs(acc).max() < 2 ** 31
    before = L.launch_count()
    got32 = BMM_S8T_S8N_S32T()(t(a), t(b))
    assert L.launch_count() - before == 1
    np.testing.assert_array_equal(got32.cpu().numpy(), acc.astype(np.int32))
    alpha = np.float32(0.37 / (K * 40.0))
    want_f = alpha * acc.astype(np.float32)  # int32 -> fp32 is exact here (|acc| < 2^24 for these K)
    assert np.abs(acc).max() < 2 ** 24
    got_f = BMM_S8T_S8N_F32T.from_scale(float(alpha), 1.0)(t(a), t(b))
    assert got_f.dtype == torch.float32
    np.testing.assert_array_equal(got_f.cpu().numpy(), want_f)
    alpha8 = np.float32(1.0 / (K * 20.0))
    want8 = np.clip(np.rint(alpha8 * acc.astype(np.float32)), -128, 127).astype(np.int8)
    got8 = BMM_S8T_S8N_S8T(float(alpha8))(t(a), t(b))
    assert got8.dtype == torch.int8 and int(np.abs(want8).max()) > 50
    np.testing.assert_array_equal(got8.cpu().numpy(), want8)


# ----------------------------------------------------------------------------- decode-sized M (weight-streaming kernel)
@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("M,N,K", [(1, 64, 256), (5, 4096, 4096), (16, 1024, 11008), (16, 12288, 4096), (9, 48, 512),
                                   (17, 256, 1088), (33, 4096, 4096), (64, 512, 8192), (40, 64, 64)])
def test_small_m_kernel_equals_tcgen05_kernel_and_oracle(dtype, M, N, K, exact_div):
    """M <= 16 launches take asq_smallm.cu (mma.sync weight stream); rows are independent, so the same rows inside
    a 96-row batch (tcgen05 kernel) must give identical bits; the smallest cases are also checked against the oracle.
    The M > 16 cases pin the hand-over back to the tcgen05 kernel."""
    rng = np.random.default_rng(M * N + K)
    td = TORCH_DT[dtype]
    x = make_x(rng, 96, K, dtype, 1.0)
    w = rng.integers(-127, 128, size=(N, K), dtype=np.int8)
    b = rng.standard_normal(N).astype(np.float32)
    cs = (rng.random(N).astype(np.float32) + 0.5) * 2e-4
    xt, wt, bt, cst = t(x, td), t(w), t(b), t(cs)
    for mode, qs in ((L.ACT_ROUND, 1.0), (L.ACT_SCALE, 0.0473), (L.ACT_PER_TOKEN, 1.0)):
        scale = 40.0 if mode == L.ACT_ROUND else 1.0
        xs = (xt.float() * scale).to(td)
        big = L.w8a8_linear(xs, wt, bt, mode, qs, 3.1e-4)
        small = L.w8a8_linear(xs[:M].contiguous(), wt, bt, mode, qs, 3.1e-4)
        assert torch.equal(small, big[:M]), f"mode {mode}: {(small != big[:M]).sum().item()} elements differ"
        big_cs = L.w8a8_linear(xs, wt, None, mode, qs, 1.0, col_scale=cst)
        small_cs = L.w8a8_linear(xs[:M].contiguous(), wt, None, mode, qs, 1.0, col_scale=cst)
        assert torch.equal(small_cs, big_cs[:M])
    q, s = L.quantize_act(xt, L.ACT_PER_TOKEN)
    assert torch.equal(L.w8a8_linear_q8(q[:M].contiguous(), wt, bt, 3.1e-4, row_scale=s[:M].contiguous(), out_dtype=td),
                       L.w8a8_linear_q8(q, wt, bt, 3.1e-4, row_scale=s, out_dtype=td)[:M])
    if N * K <= 64 * 512:
        want = O.w8a8_linear(x[:M], dtype, w, 3.1e-4, act_quant="per-token", bias=b, div_mode="exact")
        got = L.w8a8_linear(xt[:M].contiguous(), wt, bt, L.ACT_PER_TOKEN, 1.0, 3.1e-4)
        np.testing.assert_array_equal(got.float().cpu().numpy(), want)
