"""Bring-up check on a B200: unit kernels vs torch, then the whole backbone vs tests/golden/backbone.pt."""
import os
import sys
import time
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200  # noqa
from dvst_b200 import _lib
from dvst_b200.models.timesformer import VisionTransformer
from dvst_b200.testing import synth_backbone_state_dict, synth_clips

L = _lib.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
ok = True


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


# ---- LayerNorm ----
x = torch.randn(1000, 768, device=dev) * 2 + 0.3
g, b = torch.randn(768, device=dev), torch.randn(768, device=dev)
out = torch.empty(1000, 768, device=dev, dtype=torch.bfloat16)
_lib.check(L.dvst_layernorm768(_lib.ptr(x), 768, _lib.ptr(g), _lib.ptr(b), _lib.ptr(out), 768, 0, 0, 1000, 1e-6,
                               _lib.cur_stream()))
ref = torch.nn.functional.layer_norm(x, (768,), g, b, 1e-6)
print("layernorm rel", rel(out.float(), ref)); ok &= rel(out.float(), ref) < 5e-3

# ---- attention, all modes ----
def attn_ref(q, k, v):  # [S, 768] each -> [S,768]
    S = q.shape[0]
    qh, kh, vh = (t.reshape(t.shape[0], 12, 64).transpose(0, 1).float() for t in (q, k, v))
    p = torch.softmax(qh @ kh.transpose(1, 2) * 0.125, dim=-1)
    return (p @ vh).transpose(0, 1).reshape(S, 768)

for (mode, B, T, N) in [(0, 2, 16, 9), (0, 1, 3, 5), (0, 1, 30, 4), (0, 1, 60, 2), (1, 2, 3, 196), (1, 1, 2, 36), (2, 2, 2, 196),
                        (3, 2, 2, 196)]:
    Mt = B * N * T
    ncls = B * T if mode == 3 else B
    qkv = (torch.randn(Mt + ncls, 2304, device=dev) * 1.5).bfloat16()
    out = torch.zeros(Mt + ncls, 768, device=dev, dtype=torch.bfloat16)
    cls_out = torch.zeros(B, T, 768, device=dev)
    _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), mode, B, T, N, Mt,
                                _lib.cur_stream()))
    torch.cuda.synchronize()
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    refo = torch.zeros(Mt + ncls, 768, device=dev)
    refc = torch.zeros(B, T, 768, device=dev)
    rows = torch.arange(Mt, device=dev).reshape(B, N, T)
    if mode == 0:
        for bb in range(B):
            for n in range(N):
                r = rows[bb, n]
                refo[r] = attn_ref(q[r], k[r], v[r])
    elif mode in (1, 3):
        for bb in range(B):
            for t in range(T):
                c = Mt + (bb * T + t if mode == 3 else bb)
                r = torch.cat([torch.tensor([c], device=dev), rows[bb, :, t]])
                o = attn_ref(q[r], k[r], v[r])
                refo[r[1:]] = o[1:]
                if mode == 1:
                    refc[bb, t] = o[0]
                else:
                    refo[c] = o[0]
    else:
        for bb in range(B):
            r = torch.cat([torch.tensor([Mt + bb], device=dev), rows[bb].reshape(-1)])
            refo[r] = attn_ref(q[r], k[r], v[r])
    e = rel(out.float(), refo)
    msg = f"attention mode={mode} B={B} T={T} N={N} rel={e:.3e}"
    if mode == 1:
        ec = rel(cls_out, refc); msg += f" cls rel={ec:.3e}"; ok &= ec < 1e-2
        ok &= rel(out[:Mt].float(), refo[:Mt]) < 1e-2
    else:
        ok &= e < 1e-2
    print(msg, flush=True)

# ---- backbone vs golden (reference fp32 CPU outputs) ----
gold = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "backbone.pt"))
models = {}
for name, c in gold.items():
    for prec in ("bf16", "high"):
        key = (c["seed_w"], c["depth"], c["num_frames"], c["attention_type"], prec)
        if key not in models:
            m = VisionTransformer(img_size=224, num_classes=0, depth=c["depth"], qkv_bias=True,
                                  norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=c["num_frames"],
                                  attention_type=c["attention_type"], precision=prec)
            m.load_state_dict(synth_backbone_state_dict(c["seed_w"], depth=c["depth"], num_frames=c["num_frames"],
                                                        attention_type=c["attention_type"]), strict=True)
            models[key] = m.eval().to(dev)
        m = models[key]
        x = synth_clips(c["seed_x"], c["B"], c["T"], c["crop"], c["crop"]).to(dev)
        try:
            with torch.no_grad():
                f = m(x)
            torch.cuda.synchronize()
            e = rel(f.cpu(), c["feat"])
            tol = 2e-2 if prec == "bf16" else 1e-4
            print(f"backbone {name:16s} {prec:5s} rel_l2={e:.3e} max_abs={(f.cpu() - c['feat']).abs().max():.3e} "
                  f"{'OK' if e < tol else 'FAIL'}", flush=True)
            ok &= e < tol
        except Exception as ex:  # noqa
            print(f"backbone {name} {prec}: ERROR {ex}", flush=True)
            ok = False

# ---- quick timing at config 2 (B=64, T=16) ----
m = models[(0, 12, 8, "divided_space_time", "bf16")]
for (B, T) in [(8, 16), (64, 16), (64, 3), (16, 30)]:
    x = torch.randn(B, 3, T, 224, 224, device=dev)
    with torch.no_grad():
        for _ in range(2):
            m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        e0.record()
        for _ in range(3):
            m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"forward B={B} T={T}: {ms:.2f} ms  -> {B / ms * 1e3:.1f} clips/s, launches/fwd={(_lib.launch_count() - n0) // 3}",
          flush=True)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
