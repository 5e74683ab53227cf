"""A/B two builds of the library on the same box (ctypes only, no package import):
python scripts/ab_lib.py autosmoothquant_b200/libasq_old.so autosmoothquant_b200/libasq_b200.so"""
import ctypes
import sys
import torch

dev = torch.device("cuda:0")
c_vp, c_i, c_i64, c_f, c_sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
libs = []
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    lib.asq_w8a8_linear_q8.restype = c_i
    lib.asq_w8a8_linear_q8.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_f, c_vp, c_vp, c_sz, c_vp]
    lib.asq_w8a8_linear.restype = c_i
    lib.asq_w8a8_linear.argtypes = [c_vp, c_i, c_vp, c_vp, c_vp, c_i, c_i64, c_i64, c_i64, c_i, c_f, c_f, c_vp, c_vp, c_i, c_vp, c_sz, c_vp]
    lib.asq_workspace_bytes.restype = c_sz
    lib.asq_workspace_bytes.argtypes = [c_i64, c_i64]
    lib.asq_last_error.restype = ctypes.c_char_p
    libs.append((path, lib))
ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
base = (ws.data_ptr() + 1023) & ~1023


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for (M, N, K) in [(2048, 12288, 4096), (2048, 4096, 4096), (2048, 22016, 4096), (2048, 4096, 11008)]:
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    ws_ = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(3)]
    y = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    line = f"{M}x{N}x{K}:"
    for rep in range(2):
        for path, lib in libs:
            i = [0]

            def q8():
                i[0] += 1
                rc = lib.asq_w8a8_linear_q8(a.data_ptr(), None, ws_[i[0] % 3].data_ptr(), None, y.data_ptr(), 2, M, N, K, 3e-6, None,
                                            base, 200 << 20, torch.cuda.current_stream().cuda_stream)
                assert rc == 0, lib.asq_last_error()

            def fused():
                i[0] += 1
                rc = lib.asq_w8a8_linear(x.data_ptr(), 2, ws_[i[0] % 3].data_ptr(), None, y.data_ptr(), 2, M, N, K, 1, 0.05, 3e-6, None,
                                         None, 0, base, 200 << 20, torch.cuda.current_stream().cuda_stream)
                assert rc == 0

            line += f" | {path.split('/')[-1][7:-3]} q8 {timeit(q8):6.1f} fused {timeit(fused):6.1f}"
    print(line, flush=True)
