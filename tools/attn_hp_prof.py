"""Split-precision (bf16 hi / lo) spatial attention at the config-2 shape (64 clips x 16 frames x 12 heads = 12 288 problems of
197 tokens): worst error against fp64 torch on a sample of problems, and time per launch.  DVST_LIB=... times another build.
    python tools/attn_hp_prof.py [reps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, T, N = 64, 16, 196
Mt = B * N * T
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(Mt + B, 2304, device=dev, generator=g) * 1.5
hi = x.bfloat16(); lo = (x - hi.float()).bfloat16(); del x
qkv = torch.cat([hi, lo], 1).contiguous(); del hi, lo
nb = L.dvst_attention_scratch_bytes(B, T, N)
scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)
out = torch.zeros(Mt + B, 1536, device=dev, dtype=torch.bfloat16); cls_out = torch.zeros(B, T, 768, device=dev)
f = lambda: _lib.check(L.dvst_attention_hp(_lib.ptr(qkv), 4608, _lib.ptr(out), 1536, _lib.ptr(cls_out), 1, B, T, N, Mt,
                                           _lib.ptr(scratch), nb, _lib.cur_stream()))
f(); torch.cuda.synchronize()
worst = 0.0
for (b, t, h) in [(0, 0, 0), (3, 5, 7), (63, 15, 11), (17, 9, 4), (40, 1, 1)]:
    rows = torch.cat([(b * N + torch.arange(N, device=dev)) * T + t, torch.tensor([Mt + b], device=dev)])
    q, k, v = [qkv[rows, i * 768 + h * 64: i * 768 + h * 64 + 64].double() + qkv[rows, 2304 + i * 768 + h * 64: 2304 + i * 768 + h * 64 + 64].double()
               for i in range(3)]
    ref = torch.softmax(q @ k.T * 0.125, -1) @ v
    got = out[rows[:-1], h * 64:h * 64 + 64].double() + out[rows[:-1], 768 + h * 64:768 + h * 64 + 64].double()
    got = torch.cat([got, cls_out[b, t, h * 64:h * 64 + 64].double()[None]])
    worst = max(worst, ((got - ref).norm() / ref.norm()).item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): f()
e1.record(); torch.cuda.synchronize()
print(f"{os.environ.get('DVST_LIB', 'in-tree')}: {e0.elapsed_time(e1) / reps:.4f} ms per launch, worst rel-L2 vs fp64 torch {worst:.2e}, "
      f"fix-up rows {int(scratch[:4].view(torch.int32)[0])}", flush=True)
