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
