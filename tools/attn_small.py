"""Small spatial-attention launches (both split-kernel variants: N = 196 fixed geometry, N = 36 run-time geometry; modes 1 and
3) against fp32 torch - the target of compute-sanitizer memcheck / racecheck runs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda:0")
ok = True
for (mode, B, T, N) in [(1, 2, 2, 196), (3, 1, 3, 196), (1, 2, 3, 36), (1, 1, 2, 100)]:
    torch.manual_seed(N + mode)
    Mt = B * N * T
    ncls = B * T if mode == 3 else B
    qkv = (torch.randn(Mt + ncls, 2304, device=dev) * 1.5).bfloat16()
    nb = L.dvst_attention_scratch_bytes(B, T, N)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)
    out = torch.zeros(Mt + ncls, 768, device=dev, dtype=torch.bfloat16); cls_out = torch.zeros(B, T, 768, device=dev)
    _lib.check(L.dvst_attention_ws(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), mode, B, T, N, Mt,
                                   _lib.ptr(scratch), nb, _lib.cur_stream()))
    torch.cuda.synchronize()
    worst = 0.0
    for b in range(B):
        for t in range(T):
            c = Mt + (b * T + t if mode == 3 else b)
            rows = torch.cat([(b * N + torch.arange(N, device=dev)) * T + t, torch.tensor([c], device=dev)])
            for h in (0, 11):
                q, k, v = [qkv[rows, i * 768 + h * 64: i * 768 + h * 64 + 64].float() for i in range(3)]
                ref = torch.softmax(q @ k.T * 0.125, -1) @ v
                cls = cls_out[b, t, h * 64:h * 64 + 64] if mode == 1 else out[c, h * 64:h * 64 + 64].float()
                got = torch.cat([out[rows[:-1], h * 64:h * 64 + 64].float(), cls[None]])
                worst = max(worst, ((got - ref).norm() / ref.norm()).item())
    print(f"mode {mode} B {B} T {T} N {N}: worst rel-L2 {worst:.2e}")
    ok &= worst < 6e-3
print("attn small ok" if ok else "attn small MISMATCH")
sys.exit(0 if ok else 1)
