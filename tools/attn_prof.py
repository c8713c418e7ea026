"""Spatial attention at the config-2 shape (64 clips x 16 frames x 12 heads = 12 288 problems of 197 tokens): correctness of
every implementation against fp32 torch on a sample of problems, and time per launch.  Target of ncu captures.
    python tools/attn_prof.py [reps] [impl ...]     impl: 1 split tcgen05 (default), 3 first-generation one-pass, 2 two-pass"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
impls = [int(a) for a in sys.argv[2:]] or [3, 1]
B, T, N = 64, 16, 196
Mt = B * N * T
g = torch.Generator(device=dev).manual_seed(0)
qkv = (torch.randn(Mt + B, 2304, device=dev, generator=g) * 1.5).bfloat16()
nb = L.dvst_attention_scratch_bytes(B, T, N)
scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)


def reference(b, t, h):
    rows = torch.cat([(b * N + torch.arange(N, device=dev)) * T + t, torch.tensor([Mt + b], device=dev)])
    q, k, v = [qkv[rows, i * 768 + h * 64: i * 768 + h * 64 + 64].float() for i in range(3)]
    return torch.softmax(q @ k.T * 0.125, -1) @ v, rows


for impl in impls:
    L.dvst_set_attention_impl(impl)
    out = torch.zeros(Mt + B, 768, device=dev, dtype=torch.bfloat16); cls_out = torch.zeros(B, T, 768, device=dev)
    f = lambda: _lib.check(L.dvst_attention_ws(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), 1, B, T, N,
                                               Mt, _lib.ptr(scratch), nb, _lib.cur_stream()))
    f(); torch.cuda.synchronize()
    worst = 0.0
    for (b, t, h) in [(0, 0, 0), (3, 5, 7), (63, 15, 11), (17, 9, 4), (40, 1, 1)]:
        ref, rows = reference(b, t, h)
        got = torch.cat([out[rows[:-1], h * 64:h * 64 + 64].float(), cls_out[b, t, h * 64:h * 64 + 64][None]])
        worst = max(worst, ((got - ref).norm() / ref.norm()).item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    print(f"impl {impl}: {e0.elapsed_time(e1) / reps:.4f} ms per launch, worst rel-L2 vs fp32 torch {worst:.2e}, "
          f"fix-up rows {int(scratch[:4].view(torch.int32)[0])}", flush=True)
L.dvst_set_attention_impl(1)
if hasattr(L, "dvst_debug_sp2_read") or os.environ.get("DVST_LIB", "").endswith("dbg.so"):
    import ctypes
    buf = (ctypes.c_longlong * 32)()
    L.dvst_debug_sp2_read.restype = ctypes.c_int
    L.dvst_debug_sp2_read(buf)
    v = list(buf)
    n = 83.0  # problems of block 0
    print("block 0 cycles per problem:")
    print("  MMA thread : wait p_full %.0f | issue PV %.0f (of which wait v_full %.0f) | wait o_empty %.0f | issue S %.0f (of which wait "
          "qk_full %.0f)" % (v[0] / n, v[1] / n, v[5] / n, v[2] / n, v[3] / n, v[4] / n))
    for name, o in (("softmax w4 (half 0)", 8), ("softmax w8 (half 1)", 14)):
        print("  %s: wait s_full A %.0f B %.0f | reference %.0f | exp pass A %.0f B %.0f | sums+arrive %.0f" %
              ((name,) + tuple(x / n for x in v[o:o + 6])))
    print("  epilogue w12: wait p_full %.0f | wait o_full %.0f | ld O + arrive %.0f | scale/stage/store %.0f" %
          tuple(x / n for x in v[24:28]))
    tr = (ctypes.c_longlong * 128)()
    L.dvst_debug_sp2_trace(tr)
    tr = list(tr)
    t0 = min(x for x in tr if x > 0)
    names = ["mma   : p_full seen | PV issued | o_empty seen | S(u+2) issued",
             "smax0 : s_full seen | ref done | exp done | arrived p_full",
             "smax1 : s_full seen | ref done | exp done | arrived p_full",
             "epi   : p_full seen | o_full seen | O read, o_empty arrived | stored"]
    print("timeline of block 0, units 40..47 (clk since the first event):")
    for u in range(8):
        for role in range(4):
            ev = tr[(role * 8 + u) * 4:(role * 8 + u) * 4 + 4]
            print("  unit %d %s : %s" % (40 + u, names[role], "  ".join("%6d" % (x - t0) if x else "     -" for x in ev)))
