"""The plain-C restatement (oracle/asq_oracle.c) and the numpy oracle agree bit-for-bit."""
import numpy as np

from oracle import c_oracle
from oracle import w8a8_oracle as O


def test_int8_gemm_exact_including_extremes():
    rng = np.random.default_rng(1)
    for (M, N, K) in [(1, 1, 16), (7, 33, 48), (64, 96, 512), (130, 70, 1040)]:
        a = rng.integers(-128, 128, size=(M, K), dtype=np.int8)
        w = rng.integers(-128, 128, size=(N, K), dtype=np.int8)
        a[0, :] = -128
        w[0, :] = -128  # largest positive accumulator: K * 16384
        np.testing.assert_array_equal(c_oracle.i8gemm_o32(a, w), O.int8_gemm_i32(a, w))
    assert O.int8_gemm_i32(a, w)[0, 0] == 1040 * 16384


def test_per_token_quant_fp32():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((37, 96)).astype(np.float32)
    x[:, 5] *= 40
    x[3] = 0  # all-zero row: scale 0, 0/0 = NaN -> 0
    q_c, s_c = c_oracle.quant_per_token_f32(x)
    q_n, s_n = O.quantize_act_int8(x, "f32", "per-token", div_mode="exact")
    np.testing.assert_array_equal(q_c, q_n)
    np.testing.assert_array_equal(s_c, s_n)
    assert not q_n[3].any() and s_n[3] == 0


def test_round_mode_ties_to_even_and_saturation():
    x = np.array([[0.5, 1.5, 2.5, -0.5, -1.5, 126.5, 127.5, 300.0, -128.5, -129.5, -1e9, np.nan, np.inf, -np.inf, 3.49, -3.51]],
                 np.float32)
    want = np.array([[0, 2, 2, 0, -2, 126, 127, 127, -128, -128, -128, 0, 127, -128, 3, -4]], np.int8)
    np.testing.assert_array_equal(O.quantize_act_int8(x, "f32", "round")[0], want)
    np.testing.assert_array_equal(c_oracle.quant_round_f32(x), want)


def test_dequant_separately_rounded():
    rng = np.random.default_rng(3)
    acc = rng.integers(-2**30, 2**30, size=(9, 20), dtype=np.int64).astype(np.int32)
    rs = rng.random(9).astype(np.float32)
    bias = rng.standard_normal(20).astype(np.float32)
    got = c_oracle.dequant_f32(acc, rs, 0.0123, bias)
    want = O._dequant(acc, (np.float32(0.0123) * rs).reshape(-1, 1), bias, "f32")
    np.testing.assert_array_equal(got, want)


def test_e4m3_codec_three_way():
    """float8_e4m3fn: the bit-twiddling C encoder, the log2 / rint numpy encoder and torch's own cast agree on every
    finite code, on every fp16 value in range (covers all ties and the subnormal range) and on random fp32 values."""
    import torch

    codes = np.arange(256, dtype=np.uint8)
    dec_c, dec_n = c_oracle.e4m3_decode(codes), O.e4m3_decode(codes)
    dec_t = torch.from_numpy(codes).view(torch.float8_e4m3fn).float().numpy()
    finite = (codes & 0x7F) != 0x7F
    np.testing.assert_array_equal(dec_c[finite], dec_n[finite])
    np.testing.assert_array_equal(dec_c[finite], dec_t[finite])
    assert np.isnan(dec_c[~finite]).all() and np.isnan(dec_t[~finite]).all() and dec_c[0x7E] == 448.0
    # round trip of every finite code
    np.testing.assert_array_equal(c_oracle.e4m3_encode(dec_c[finite]), codes[finite])
    np.testing.assert_array_equal(O.e4m3_encode(dec_n[finite]), codes[finite])
    # every fp16 value with |v| <= 448 (ties at every quantum boundary included), then random fp32
    half = np.arange(1 << 16, dtype=np.uint16).view(np.float16).astype(np.float32)
    half = half[np.isfinite(half) & (np.abs(half) <= 448.0)]
    rng = np.random.default_rng(0)
    rand = np.concatenate([rng.uniform(-448, 448, 200000), rng.standard_normal(200000) * 0.01, rng.standard_normal(100000) * 2.0 ** -8]).astype(np.float32)
    for v in (half, rand):
        want = torch.from_numpy(v).to(torch.float8_e4m3fn).view(torch.uint8).numpy()
        got_c, got_n = c_oracle.e4m3_encode(v), O.e4m3_encode(v)
        # -0.0 and values that round to zero keep their sign bit in all three
        np.testing.assert_array_equal(got_c, want)
        np.testing.assert_array_equal(got_n, want)


def test_fp8_per_token_quant_and_exact_linear():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((29, 80)).astype(np.float32)
    x[:, 7] *= 60
    x[2] = 0  # all-zero token: scale 0, 0/0 -> NaN codes
    q_c, s_c = c_oracle.fp8_quant_per_token_f32(x)
    q_n, s_n = O.quantize_act_fp8(x, "f32", "per-token", div_mode="exact")
    np.testing.assert_array_equal(s_c, s_n)
    nan = (q_n & 0x7F) == 0x7F
    np.testing.assert_array_equal((q_c & 0x7F) == 0x7F, nan)
    np.testing.assert_array_equal(q_c[~nan], q_n[~nan])
    assert nan[2].all() and not nan[[0, 1, 3]].any()
    rows = [i for i in range(29) if i != 2]
    w = O.e4m3_encode(np.clip(rng.standard_normal((24, 80)).astype(np.float32) * 30, -448, 448))
    bias = rng.standard_normal(24).astype(np.float32)
    w_scale = float(np.float32(0.013))  # the weight scale is an fp32 buffer on both sides
    y_c = c_oracle.fp8_linear_f64(q_c[rows], w, s_c[rows], w_scale, bias)
    y_n = O.fp8_linear_exact(q_n[rows], w, s_n[rows], w_scale, bias)
    np.testing.assert_allclose(y_c, y_n, rtol=1e-12, atol=1e-12 * np.abs(y_n).max())
