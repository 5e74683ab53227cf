"""GPU bring-up checks run under gpurun (not part of the pytest suite).

usage: python scripts/bringup.py {gemm|quant|linear|fp8|perf} ...
Each stage prints PASS/FAIL lines; run stages under `timeout -s KILL` so a hung kernel cannot
hold the box.
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from autosmoothquant_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda:0")


def exact_i32(a, w):
    return (a.double() @ w.double().t()).to(torch.int32)


def report(name, ok, extra=""):
    print(f"{'PASS' if ok else 'FAIL'} {name} {extra}", flush=True)


def stage_gemm():
    print("device supported:", L.load().asq_device_supported(), torch.cuda.get_device_name(0), flush=True)
    shapes = [
        (128, 256, 128), (128, 256, 512), (128, 64, 128), (128, 128, 256), (256, 512, 4096),
        (100, 300, 144), (1, 4096, 4096), (129, 257, 272), (2048, 4096, 4096), (2048, 11008, 4096),
        (2048, 4096, 11008), (333, 1000, 1008), (4096, 768, 3072),
    ]
    for (M, N, K) in shapes:
        g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8, generator=g).to(dev)
        w = torch.randint(-128, 128, (N, K), dtype=torch.int8, generator=g).to(dev)
        out = torch.full((M, N), -77, dtype=torch.int32, device=dev)
        try:
            L.i8gemm_o32(a, w, out)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            report(f"gemm {M}x{N}x{K}", False, f"exception {e}")
            continue
        ref = exact_i32(a, w)
        bad = (out != ref)
        nbad = int(bad.sum())
        extra = ""
        if nbad:
            idx = bad.nonzero()[:6].tolist()
            rows_bad = bad.any(dim=1).sum().item()
            cols_bad = bad.any(dim=0).sum().item()
            extra = f"nbad={nbad}/{M * N} rows_bad={rows_bad} cols_bad={cols_bad} first={idx} " \
                    f"got={[int(out[i, j]) for i, j in idx]} want={[int(ref[i, j]) for i, j in idx]}"
        report(f"gemm {M}x{N}x{K}", nbad == 0, extra)


def ref_quant(x, mode, qs, recip, fp8):
    """torch-eager restatement of the reference prologue on the same device as x."""
    qmax = 448.0 if fp8 else 127.0
    if mode == L.ACT_PER_TOKEN:
        amax = x.abs().max(dim=-1, keepdim=True)[0]
        if recip:
            s = (amax.float() * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(qmax, dtype=torch.float32)).item()).to(x.dtype).float()
        else:
            s = (amax.float() / qmax).to(x.dtype).float()
        v = x.float() / s
    elif mode == L.ACT_SCALE:
        if recip:
            inv = (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(qs, dtype=torch.float32)).item()
            v = (x.float() * inv).to(x.dtype).float()
        else:
            v = (x.float() / torch.tensor(qs, dtype=torch.float32)).to(x.dtype).float()
        s = None
    else:
        v = x.float()
        s = None
    if fp8:
        q = v.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)
    else:
        q = v.round().clamp(-128, 127).to(torch.int8)
    return q, (s.flatten() if s is not None else None)


def make_x(M, K, dtype, seed, scale=1.0, outliers=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(M, K, generator=g) * scale
    if outliers:
        cols = torch.randint(0, K, (max(1, K // 512),), generator=g)
        x[:, cols] *= 30.0
    return x.to(dtype).to(dev)


def stage_quant():
    for dtype in (torch.bfloat16, torch.float16, torch.float32):
        for (M, K) in [(4, 128), (130, 4096), (2048, 4096), (257, 11008)]:
            for mode, name in ((L.ACT_ROUND, "round"), (L.ACT_SCALE, "scale"), (L.ACT_PER_TOKEN, "token")):
                for recip in (True, False):
                    for fp8 in (False, True):
                        if fp8 and mode == L.ACT_ROUND:
                            continue
                        sc = 40.0 if mode == L.ACT_ROUND else 1.0
                        x = make_x(M, K, dtype, M + K, sc)
                        if M > 2:
                            x[1].zero_()  # all-zero row
                        qs = 0.0473
                        q, rs = L.quantize_act(x, mode, qs, fp8=fp8, div_mode=L.DIV_RECIPROCAL if recip else L.DIV_EXACT)
                        torch.cuda.synchronize()
                        # torch on CUDA divides by scalars through the reciprocal; the exact mode is the CPU behaviour
                        qr, sr = ref_quant(x if recip else x.cpu(), mode, qs, recip, fp8)
                        qr = qr.to(dev)
                        sr = sr.to(dev) if sr is not None else None
                        if fp8:
                            a, b = q.view(torch.uint8), qr.view(torch.uint8)
                            # all-zero rows: reference gives NaN bytes (0/0); compare bitwise anyway
                            nbad = int((a != b).sum())
                        else:
                            nbad = int((q != qr).sum())
                        sbad = 0 if rs is None else int((rs != sr).sum())
                        report(f"quant {str(dtype)[6:]} {M}x{K} {name} {'recip' if recip else 'exact'} {'fp8' if fp8 else 'i8'}",
                               nbad == 0 and sbad == 0, f"nbad={nbad} sbad={sbad}")


def ref_linear(x, w, bias, mode, qs, ds, recip, col_scale=None):
    q, s = ref_quant(x, mode, qs, recip, False)
    acc = exact_i32(q, w)
    if col_scale is not None:
        f = col_scale.view(1, -1)
        if s is not None:
            f = f * s.view(-1, 1)
        y = f * acc
    elif s is not None:
        y = (ds * s.view(-1, 1)) * acc
    else:
        y = ds * acc
    if bias is not None:
        y = y + bias
    return y.to(x.dtype)


def stage_linear():
    for dtype in (torch.bfloat16, torch.float16, torch.float32):
        for (M, N, K) in [(128, 256, 128), (200, 1000, 528), (2048, 4096, 4096), (77, 11008, 4096), (2048, 4096, 11008)]:
            for mode, name in ((L.ACT_ROUND, "round"), (L.ACT_SCALE, "scale"), (L.ACT_PER_TOKEN, "token")):
                for use_bias in (False, True):
                    sc = 40.0 if mode == L.ACT_ROUND else 1.0
                    x = make_x(M, K, dtype, M + N + K, sc)
                    g = torch.Generator(device="cpu").manual_seed(N + K)
                    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g).to(dev)
                    bias = torch.randn(N, generator=g).to(dev) if use_bias else None
                    qs, ds = 0.0473, 0.00321
                    for rep in range(2):  # second call re-uses the workspace counters
                        y = L.w8a8_linear(x, w, bias, mode, qs, ds, div_mode=L.DIV_RECIPROCAL)
                    torch.cuda.synchronize()
                    yr = ref_linear(x, w, bias, mode, qs, ds, True)
                    nbad = int((y != yr).sum()) if not torch.isnan(yr).any() else int((y.float() - yr.float()).abs().nan_to_num().gt(0).sum())
                    md = float((y.float() - yr.float()).abs().max())
                    report(f"linear {str(dtype)[6:]} {M}x{N}x{K} {name} bias={use_bias}", nbad == 0, f"nbad={nbad} maxdiff={md:.3e}")
    # col_scale (QKV) variant
    M, N, K = 300, 768, 256
    x = make_x(M, K, torch.bfloat16, 5)
    g = torch.Generator(device="cpu").manual_seed(9)
    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g).to(dev)
    cs = torch.cat([torch.full((256,), v) for v in (0.001, 0.002, 0.003)]).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    for mode, name in ((L.ACT_ROUND, "round"), (L.ACT_PER_TOKEN, "token")):
        y = L.w8a8_linear(x, w, bias, mode, 1.0, 1.0, col_scale=cs)
        yr = ref_linear(x, w, bias, mode, 1.0, 1.0, True, col_scale=cs)
        report(f"linear colscale {name}", bool((y == yr).all()), f"maxdiff={float((y.float() - yr.float()).abs().max()):.3e}")


def stage_fp8():
    for dtype in (torch.bfloat16, torch.float32):
        for (M, N, K) in [(128, 256, 128), (200, 1000, 528), (2048, 4096, 4096)]:
            for mode, name in ((L.ACT_SCALE, "static"), (L.ACT_PER_TOKEN, "token")):
                x = make_x(M, K, dtype, M + N + K)
                g = torch.Generator(device="cpu").manual_seed(N + K)
                wf = torch.randn(N, K, generator=g) * 0.02
                ws = float(wf.abs().max() / 448.0)
                w = (wf / ws).clamp(-448, 448).to(torch.float8_e4m3fn).to(dev)
                in_scale = float(x.float().abs().max() / 448.0)
                y = L.fp8_linear(x, w, None, mode, in_scale, ws)
                torch.cuda.synchronize()
                q, s = ref_quant(x, mode, in_scale, True, True)
                acc = q.double() @ w.double().t()
                if s is not None:
                    yr = acc * (s.double().view(-1, 1) * ws)
                else:
                    yr = acc * (ws * in_scale)
                err = (y.double() - yr).abs().max().item()
                scale = yr.abs().max().item()
                report(f"fp8 {str(dtype)[6:]} {M}x{N}x{K} {name}", err <= scale * 2 ** -7, f"maxerr={err:.3e} ref_absmax={scale:.3e}")


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def stage_perf():
    shapes = [(2048, 4096, 4096), (2048, 12288, 4096), (2048, 11008, 4096), (2048, 22016, 4096), (2048, 4096, 11008),
              (8192, 8192, 8192), (16, 4096, 4096)]
    for (M, N, K) in shapes:
        a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
        w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
        out = torch.empty((M, N), dtype=torch.int32, device=dev)
        x = (torch.randn(M, K, device=dev) * 40).to(torch.bfloat16)
        t_gemm = timeit(lambda: L.i8gemm_o32(a, w, out))
        ops = 2.0 * M * N * K
        line = f"perf {M}x{N}x{K}: i8gemm_o32 {t_gemm * 1e6:8.1f} us {ops / t_gemm / 1e12:7.1f} TOPS"
        for mode, name in ((L.ACT_ROUND, "round"), (L.ACT_SCALE, "scale"), (L.ACT_PER_TOKEN, "token")):
            t = timeit(lambda: L.w8a8_linear(x, w, None, mode, 0.05, 0.003))
            line += f" | fused-{name} {t * 1e6:8.1f} us {ops / t / 1e12:7.1f} TOPS"
        try:
            t_mm = timeit(lambda: torch._int_mm(a, w.t())) if M > 16 else float("nan")
            line += f" | torch._int_mm {t_mm * 1e6:8.1f} us {ops / t_mm / 1e12:7.1f} TOPS"
        except Exception as e:  # noqa: BLE001
            line += f" | torch._int_mm failed: {type(e).__name__}"
        print(line, flush=True)


if __name__ == "__main__":
    t0 = time.time()
    for st in sys.argv[1:]:
        globals()["stage_" + st]()
    print(f"done in {time.time() - t0:.1f}s", flush=True)
