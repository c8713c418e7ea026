"""The K = 768 deferred-LayerNorm producer GEMM (proj shape) timed alone and inside a power-capped stream of tensor-bound
GEMMs (one fc1-sized GEMM between every two producer launches, as in the step).  usage: python tools/producer_prof.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200  # noqa: F401,E402
from dvst_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
M, K = 200704, 768
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
torch.manual_seed(0)
A = torch.randn(M, K, device=dev).bfloat16()
W = (torch.randn(768, K, device=dev) * 0.05).bfloat16()
bias = torch.randn(768, device=dev)
x = torch.randn(M, 768, device=dev)
xb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
stats = torch.empty(M, 6, 2, device=dev)
W1 = (torch.randn(3072, 768, device=dev) * 0.05).bfloat16()
hid = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
st = _lib.cur_stream()
BYTES = M * (2.0 * K + 10.0 * 768) + 2.0 * 768 * K


def producer():
    _lib.check(L.dvst_gemm_bf16_ln(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, 768, K, None, None, 0.0, _lib.ptr(x), 768,
                                   _lib.ptr(xb), 768, _lib.ptr(stats), None, None, 0, 0, st))


def heavy():
    _lib.check(L.dvst_gemm_bf16(_lib.ptr(xb), 768, _lib.ptr(W1), 768, None, M, 3072, 768, 0, None, 0, _lib.ptr(hid), 3072, 0, 0,
                                1, st))


def run(with_heavy, reps):
    for _ in range(5):
        if with_heavy:
            heavy()
        producer()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in ev:
        if with_heavy:
            heavy()
            heavy()
        e0.record()
        producer()
        e1.record()
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return sum(t) / len(t), t[len(t) // 2]


for wh in (False, True):
    avg, med = run(wh, n * (4 if wh else 1))
    print(f"producer K=768 {'between fc1+GELU GEMMs (power-capped)' if wh else 'alone'}: avg {avg:.4f} ms, median {med:.4f} ms"
          f" = {BYTES / med / 1e6:.0f} GB/s", flush=True)
