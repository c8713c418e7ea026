"""Launch each config-2 GEMM shape a few times (target of the ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for (M, N, K, kind, act, name) in [(200704, 2304, 768, 0, 0, "qkv"), (200704, 768, 768, 1, 0, "proj+res"),
                                   (200768, 3072, 768, 0, 1, "fc1+gelu"), (200768, 768, 3072, 1, 0, "fc2+res")]:
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev) if kind == 1 else None
    out = res if kind == 1 else torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, 0, _lib.ptr(res), N,
                                    _lib.ptr(out), N, kind, 0, act, _lib.cur_stream()))
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / reps, "ms", 2.0 * M * N * K * reps / e0.elapsed_time(e1) / 1e9, "TFLOP/s", flush=True)
