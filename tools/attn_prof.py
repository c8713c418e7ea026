"""Launch the config-2 attention shapes (spatial tcgen05 kernel, temporal kernel) a few times - target of ncu captures."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, T, N = 64, 16, 196
Mt = B * N * T
qkv = (torch.randn(Mt + B, 2304, device=dev) * 1.5).bfloat16()
out = torch.zeros(Mt + B, 768, device=dev, dtype=torch.bfloat16); cls_out = torch.zeros(B, T, 768, device=dev)
for mode in (1, 0):
    f = lambda: _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), mode, B, T, N,
                                            Mt, _lib.cur_stream()))
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    print("mode", mode, e0.elapsed_time(e1) / reps, "ms", flush=True)
