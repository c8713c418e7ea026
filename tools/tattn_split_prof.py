"""Split-precision (bf16 hi / lo) temporal attention at the config-2 shape (64 clips x 196 patches = 12 544 sequences of T tokens,
12 heads): time per launch and achieved bandwidth of attention_kernel<1, true>.  (Round 2: a contiguous-sequence variant - one CTA
per sequence and 6 heads, coalesced 768-byte slices - measured 0.631 vs 0.676 ms at T = 16, 0.89 vs 0.61 ms at T = 30 (one CTA per SM),
equal at T = 3 / 8: the kernel is bound by the load -> mma chain -> store latency of its small CTAs, not by coalescing; not kept.)
    python tools/tattn_split_prof.py [reps] [T ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
Ts = [int(a) for a in sys.argv[2:]] or [16, 30, 3]
for T in Ts:
    B, N = (64, 196) if T <= 16 else (32, 196)
    Mt = B * N * T
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(Mt + B, 2304, device=dev, generator=g) * 1.5
    hi = x.bfloat16(); qkv = torch.cat([hi, (x - hi.float()).bfloat16()], 1).contiguous(); del x, hi
    out = torch.zeros(Mt + B, 1536, device=dev, dtype=torch.bfloat16)
    gb = Mt * (4608 + 1536) * 2 / 1e9
    for impl in (1,):
        f = lambda: _lib.check(L.dvst_attention_hp(_lib.ptr(qkv), 4608, _lib.ptr(out), 1536, None, 0, B, T, N, Mt, None, 0,
                                                   _lib.cur_stream()))
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"T={T} B={B} impl {impl}: {ms:.4f} ms per launch, {gb / ms * 1e3:.0f} GB/s of {gb:.2f} GB", flush=True)
    del qkv, out
