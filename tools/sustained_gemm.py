"""Sustained (power-capped) throughput of GEMM epilogue variants at cfg2 shapes: each case loops for ~1.5 s, so the number
reflects energy per unit of work, not burst clocks.  torch.matmul (cuBLAS) of the same shape beside it.
usage: python tools/sustained_gemm.py [seconds]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200  # noqa: F401,E402
from dvst_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
only = sys.argv[2] if len(sys.argv) > 2 else None     # substring filter on the case name
M = 200704
st = _lib.cur_stream()


def sustained(f, flops):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    # calibrate
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    n = max(10, int(secs * 1e3 / e0.elapsed_time(e1)))
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return ms, flops / ms / 1e9


def case(name, N, K, mode):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    flops = 2.0 * M * N * K
    if mode == "cublas":
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        f = lambda: torch.matmul(A, W.t(), out=out)
    elif mode in ("plain", "gelu"):
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        f = lambda: _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, 0, None, 0, _lib.ptr(out),
                                                N, 0, 0, 1 if mode == "gelu" else 0, st))
    elif mode in ("ln", "ln_gelu"):
        stats = torch.rand(M, 6, 2, device=dev) * 50 + 150
        sf = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        f = lambda: _lib.check(L.dvst_gemm_bf16_ln(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, _lib.ptr(stats),
                                                   _lib.ptr(sf), 1e-6, None, 0, None, 0, None, None, _lib.ptr(out), N,
                                                   1 if mode == "ln_gelu" else 0, st))
    elif mode == "f32_reduce":
        x = torch.zeros(M, N, device=dev)
        f = lambda: _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, 0, _lib.ptr(x), N, _lib.ptr(x),
                                                N, 1, 0, 0, st))
    elif mode == "producer":
        x = torch.zeros(M, N, device=dev)
        xb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        stats = torch.empty(M, 6, 2, device=dev)
        f = lambda: _lib.check(L.dvst_gemm_bf16_ln(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, None, None, 0.0, _lib.ptr(x),
                                                   N, _lib.ptr(xb), N, _lib.ptr(stats), None, None, 0, 0, st))
    ms, tf = sustained(f, flops)
    print(f"{name:44s} N={N:5d} K={K:5d}: {ms:.4f} ms  {tf:7.0f} TFLOP/s sustained", flush=True)


for name, N, K, mode in [("fc1 cuBLAS (no epilogue)", 3072, 768, "cublas"), ("fc1 plain bf16 + bias", 3072, 768, "plain"),
                         ("fc1 + GELU", 3072, 768, "gelu"), ("fc1 deferred-LN", 3072, 768, "ln"),
                         ("fc1 deferred-LN + GELU (the step's form)", 3072, 768, "ln_gelu"),
                         ("qkv cuBLAS", 2304, 768, "cublas"), ("qkv plain bf16 + bias", 2304, 768, "plain"),
                         ("qkv deferred-LN (the step's form)", 2304, 768, "ln"),
                         ("fc2 cuBLAS", 768, 3072, "cublas"), ("fc2 fp32 reduce-add residual", 768, 3072, "f32_reduce"),
                         ("fc2 producer (the step's form)", 768, 3072, "producer")]:
    if only is None or only in name:
        case(name, N, K, mode)
