"""Fused qkv-GEMM + temporal attention (csrc/gemm_tattn_sm100.cu) against the two-kernel form, timed with CUDA events.
usage: python tools/tattn_prof.py [T n_seq]..."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dvst_b200  # noqa: F401,E402
from dvst_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(T, n_seq):
    M = T * n_seq
    torch.manual_seed(0)
    x = torch.randn(M, 768, device=dev)
    W = torch.randn(2304, 768, device=dev) * 0.05
    b = torch.randn(2304, device=dev) * 0.1
    xb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(M, 6, 2, device=dev)
    _lib.check(L.dvst_rowstats768(_lib.ptr(x), 768, _lib.ptr(xb), 768, _lib.ptr(stats), None, M, _lib.cur_stream()))
    wf = W.bfloat16().contiguous()
    sf = wf.double().sum(1).float().contiguous()
    hp = torch.arange(2304, device=dev).view(3, 12, 64).permute(1, 0, 2).reshape(-1)
    wfp, sfp, bp = wf[hp].contiguous(), sf[hp].contiguous(), b[hp].contiguous()
    qkv = torch.empty(M, 2304, device=dev, dtype=torch.bfloat16)
    o2 = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    o1 = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    st = _lib.cur_stream()

    def gemm():
        _lib.check(L.dvst_gemm_bf16_ln(_lib.ptr(xb), 768, _lib.ptr(wf), 768, _lib.ptr(b), M, 2304, 768, _lib.ptr(stats),
                                       _lib.ptr(sf), 1e-6, None, 0, None, 0, None, None, _lib.ptr(qkv), 2304, 0, st))

    def att():
        _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(o2), 768, None, 0, n_seq, T, 1, M, st))

    def fused():
        _lib.check(L.dvst_gemm_ln_temporal_attention(_lib.ptr(xb), 768, _lib.ptr(wfp), _lib.ptr(bp), _lib.ptr(sfp),
                                                     _lib.ptr(stats), 1e-6, M, T, 0, _lib.ptr(o1), 768, st))

    tg, ta, tf = timeit(gemm), timeit(att), timeit(fused)
    gemm(); att(); fused()
    torch.cuda.synchronize()
    same = bool(torch.equal(o1, o2))
    rel = ((o1.float() - o2.float()).norm() / o2.float().norm()).item()
    fl = 2.0 * M * 2304 * 768
    print(f"T={T:3d} n_seq={n_seq:6d} M={M:7d}: qkv GEMM {tg:.4f} ms ({fl / tg / 1e9:.0f} TF/s) + temporal attention {ta:.4f} ms"
          f" = {tg + ta:.4f} | fused {tf:.4f} ms ({fl / tf / 1e9:.0f} TF/s) | bits equal {same}, rel {rel:.2e}", flush=True)


if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:]]
    cases = list(zip(args[0::2], args[1::2])) or [(16, 12544), (8, 50176), (4, 50176), (32, 6272), (30, 6272), (3, 25088),
                                                  (16, 196)]
    for T, n in cases:
        run(T, n)
