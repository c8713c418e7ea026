"""Spatial / space-only attention through the C ABI: tcgen05 kernel vs the mma.sync kernel vs a torch fp32 reference,
plus timing of both implementations at the config-2 shape (B=64, T=16, N=196)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")


def ref_attn(q, k, v):
    q, k, v = (x.float().reshape(-1, 12, 64).transpose(0, 1) for x in (q, k, v))
    p = torch.softmax(q @ k.transpose(1, 2) * 0.125, dim=-1)
    return (p @ v).transpose(0, 1).reshape(-1, 768)


def run(mode, B, T, N, impl, qkv):
    Mt = B * N * T
    ncls = B * T if mode == 3 else B
    out = torch.zeros(Mt + ncls, 768, device=dev, dtype=torch.bfloat16)
    cls_out = torch.zeros(B, T, 768, device=dev)
    L.dvst_set_attention_impl(impl)
    _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), mode, B, T, N, Mt,
                                _lib.cur_stream()))
    torch.cuda.synchronize()
    return out, cls_out


ok = True
for (mode, B, T, N) in [(1, 1, 1, 196), (1, 2, 3, 196), (3, 2, 2, 196), (1, 2, 2, 36), (3, 1, 3, 36), (1, 3, 5, 128),
                        (1, 2, 2, 127), (1, 1, 2, 255), (1, 40, 4, 196)]:
    torch.manual_seed(mode * 1000 + N + T)
    Mt = B * N * T
    ncls = B * T if mode == 3 else B
    qkv = (torch.randn(Mt + ncls, 2304, device=dev) * 1.5).bfloat16()
    o1, c1 = run(mode, B, T, N, 1, qkv)
    o0, c0 = run(mode, B, T, N, 0, qkv)
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    rows = torch.arange(Mt, device=dev).reshape(B, N, T)
    refo = torch.zeros(Mt + ncls, 768, device=dev); refc = torch.zeros(B, T, 768, device=dev)
    for bb in range(B):
        for t in range(T):
            c = Mt + (bb * T + t if mode == 3 else bb)
            r = torch.cat([torch.tensor([c], device=dev), rows[bb, :, t]])
            o = ref_attn(q[r], k[r], v[r])
            refo[r[1:]] = o[1:]
            if mode == 1: refc[bb, t] = o[0]
            else: refo[c] = o[0]
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()
    e_new, e_old = rel(o1, refo), rel(o0, refo)
    ec_new = rel(c1, refc) if mode == 1 else 0.0
    nan = bool(torch.isnan(o1.float()).any() or torch.isnan(c1).any())
    good = e_new < 5e-3 and ec_new < 5e-3 and not nan
    if mode == 1: good &= float(o1[Mt:].float().abs().sum()) == 0
    ok &= good
    print(f"mode {mode} B={B} T={T} N={N}: tcgen05 rel {e_new:.3e} (cls {ec_new:.3e}) | mma.sync rel {e_old:.3e} | "
          f"new-vs-old {rel(o1, o0):.3e} nan={nan} {'OK' if good else 'FAIL'}", flush=True)

B, T, N = 64, 16, 196
Mt = B * N * T
qkv = (torch.randn(Mt + B, 2304, device=dev) * 1.5).bfloat16()
out = torch.zeros(Mt + B, 768, device=dev, dtype=torch.bfloat16); cls_out = torch.zeros(B, T, 768, device=dev)
for impl in (0, 1):
    L.dvst_set_attention_impl(impl)
    f = lambda: _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), 1, B, T, N, Mt,
                                            _lib.cur_stream()))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cfg2 spatial attention impl={impl}: {ms:.3f} ms  ({(Mt + B * T) * 6144 / ms / 1e6:.0f} GB/s algorithmic, "
          f"{4.0 * B * T * 12 * 197 * 197 * 64 / ms / 1e9:.0f} TFLOP/s)", flush=True)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
