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

# deferred-LayerNorm forms (dvst_gemm_bf16_ln): producer = in-place residual + bf16 copy + row stats, consumer = kLN epilogue
for (M, K, name) in [(200768, 3072, "fc2 producer (res + xb + stats)"), (200704, 768, "proj producer (res + xb + stats)")]:
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(768, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(768, device=dev); x = torch.randn(M, 768, device=dev)
    xb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16); stats = torch.empty(M, 6, 2, device=dev)
    f = lambda: _lib.check(L.dvst_gemm_bf16_ln(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, 768, K, None, None, 0.0,
                                               _lib.ptr(x), 768, _lib.ptr(xb), 768, _lib.ptr(stats), None, None, 0, 0,
                                               _lib.cur_stream()))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:36s} M={M} N=768 K={K}: {ms:.3f} ms = {2.0 * M * 768 * K / ms / 1e9:.0f} TFLOP/s", flush=True)
for (M, N, act, name) in [(200704, 2304, 0, "qkv consumer (kLN)"), (200768, 3072, 1, "fc1+gelu consumer (kLN)")]:
    xb = torch.randn(M, 768, device=dev).bfloat16(); W = (torch.randn(N, 768, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev); sf = torch.randn(N, device=dev)
    stats = torch.rand(M, 6, 2, device=dev) * 50 + 150
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    f = lambda: _lib.check(L.dvst_gemm_bf16_ln(_lib.ptr(xb), 768, _lib.ptr(W), 768, _lib.ptr(bias), M, N, 768,
                                               _lib.ptr(stats), _lib.ptr(sf), 1e-6, None, 0, None, 0, None, None, _lib.ptr(out),
                                               N, act, _lib.cur_stream()))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:36s} M={M} N={N} K=768: {ms:.3f} ms = {2.0 * M * N * 768 / ms / 1e9:.0f} TFLOP/s", flush=True)
