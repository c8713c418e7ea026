"""Bring-up check of the tcgen05 GEMM through the C ABI (run on a B200 via gpurun)."""
import ctypes
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200  # noqa: E402
from dvst_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")


def gemm(A, W, bias=None, residual=None, out_kind=0, act=0, split=0, K=None):
    M = A.shape[0]
    N = W.shape[0]
    K = K or A.shape[1]
    if out_kind == 1:
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        ldo, lo_off = N, 0
    elif out_kind == 0:
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ldo, lo_off = N, 0
    else:
        out = torch.empty(M, 2 * N, device=dev, dtype=torch.bfloat16)
        ldo, lo_off = 2 * N, N
    _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(bias), M, N, K, split,
                                _lib.ptr(residual), residual.stride(0) if residual is not None else 0,
                                _lib.ptr(out), ldo, out_kind, lo_off, act, _lib.cur_stream()))
    return out


def report(name, got, ref):
    got = got.float()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()
    print(f"{name:50s} max_abs={err:.3e} rel_l2={rel:.3e} {'OK' if rel < 1e-2 else 'FAIL'}", flush=True)
    return rel < 1e-2


torch.manual_seed(0)
ok = True
CG = int(sys.argv[1]) if len(sys.argv) > 1 else 2
L.dvst_set_gemm_cta_group(CG)
print("cta_group", CG, flush=True)

# 1. identity-weight probe: out[m, n] = A[m, n] for n < K
M, N, K = 128, 256, 64
A = torch.randn(M, K, device=dev).bfloat16()
W = torch.zeros(N, K, device=dev).bfloat16()
W[torch.arange(K), torch.arange(K)] = 1
out = gemm(A, W, out_kind=1)
torch.cuda.synchronize()
ref = A.float() @ W.float().t()
if not report("identity probe 128x256x64", out, ref):
    ok = False
    print("out[0,:16] ", out[0, :16].tolist())
    print("ref[0,:16] ", ref[0, :16].tolist())
    print("out[:16,0] ", out[:16, 0].tolist())
    print("ref[:16,0] ", ref[:16, 0].tolist())
    nz = (out != 0).nonzero()
    print("nonzero count", nz.shape[0], "expected", (ref != 0).sum().item())

for (M, N, K) in [(128, 256, 64), (128, 256, 768), (256, 512, 128), (1000, 768, 768), (333, 2304, 768),
                  (4099, 3072, 768), (4099, 768, 3072), (40000, 768, 768), (37999, 2304, 768)]:
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    ref = A.float() @ W.float().t() + bias
    out = gemm(A, W, bias=bias, out_kind=1)
    ok &= report(f"f32 out  M={M} N={N} K={K}", out, ref)
    out = gemm(A, W, bias=bias, out_kind=0)
    ok &= report(f"bf16 out M={M} N={N} K={K}", out, ref)
    out = gemm(A, W, bias=bias, out_kind=0, act=1)
    ok &= report(f"gelu     M={M} N={N} K={K}", out, torch.nn.functional.gelu(ref))
    res = torch.randn(M, N, device=dev)
    x = res.clone()
    _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, 0, _lib.ptr(x), N,
                                _lib.ptr(x), N, 1, 0, 0, _lib.cur_stream()))
    ok &= report(f"residual in-place M={M} N={N} K={K}", x, ref + res)

# split precision: compare against fp64 product of fp32 inputs
M, N, K = 777, 768, 768
Af = torch.randn(M, K, device=dev)
Wf = torch.randn(N, K, device=dev) * 0.05


def hilo(t):
    hi = t.bfloat16()
    lo = (t - hi.float()).bfloat16()
    return torch.cat([hi, lo], dim=1).contiguous()


ref64 = (Af.double() @ Wf.double().t())
out = gemm(hilo(Af), hilo(Wf), out_kind=1, split=1, K=K)
e_split = ((out.double() - ref64).norm() / ref64.norm()).item()
out_plain = gemm(Af.bfloat16(), Wf.bfloat16(), out_kind=1)
e_plain = ((out_plain.double() - ref64).norm() / ref64.norm()).item()
print(f"split precision rel err {e_split:.3e} (plain bf16 {e_plain:.3e})", flush=True)
ok &= e_split < 1e-4
out2 = gemm(hilo(Af), hilo(Wf), out_kind=2, split=1, K=K)
rec = out2[:, :N].float() + out2[:, N:].float()
print(f"hi/lo output rel err {((rec.double() - ref64).norm() / ref64.norm()).item():.3e}", flush=True)

# timing at config-2 sizes
for (M, N, K, kind, act, name) in [(200704, 2304, 768, 0, 0, "qkv"), (200704, 768, 768, 1, 0, "proj+res"),
                                   (200768, 3072, 768, 0, 1, "fc1+gelu"), (200768, 768, 3072, 1, 0, "fc2+res")]:
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev) if kind == 1 else None
    out = res if kind == 1 else torch.empty(M, N, device=dev, dtype=torch.bfloat16)

    def run():
        _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, 0, _lib.ptr(res), N,
                                    _lib.ptr(out), N, kind, 0, act, _lib.cur_stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS comparison
    for _ in range(3):
        torch.matmul(A, W.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        torch.matmul(A, W.t())
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print(f"time {name:10s} M={M} N={N} K={K}: {ms:.3f} ms = {tf:.1f} TFLOP/s | cuBLAS {ms2:.3f} ms = "
          f"{2.0 * M * N * K / ms2 / 1e9:.1f} TFLOP/s", flush=True)

print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
