"""Timing of GEMM variants at config-2 shapes (experiments): python tools/gemm_shapes.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
cases = [(200768, 768, 3072, 1, 0, "fc2 fp32 reduce-add"), (200768, 768, 3072, 0, 0, "fc2 shape, bf16 store"),
         (200768, 768, 3072, 1, 0, "fc2 fp32 plain store (no residual)"),
         (200704, 768, 768, 1, 0, "proj fp32 reduce-add"), (200704, 768, 768, 0, 0, "proj shape, bf16 store"),
         (200704, 2304, 768, 0, 0, "qkv"), (200768, 3072, 768, 0, 1, "fc1+gelu"), (200768, 3072, 768, 0, 0, "fc1 no gelu")]
for (M, N, K, kind, act, name) in cases:
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev) if (kind == 1 and "no residual" not in name) else None
    out = res if res is not None else torch.empty(M, N, device=dev, dtype=torch.float32 if kind == 1 else torch.bfloat16)
    f = lambda: _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, 0, _lib.ptr(res), N,
                                            _lib.ptr(out), N, kind, 0, act, _lib.cur_stream()))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:36s} M={M} N={N} K={K}: {ms:.3f} ms = {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
