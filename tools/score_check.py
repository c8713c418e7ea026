"""Bring-up check on a B200: DINO head / losses / whole-video scoring vs tests/golden (reference outputs)."""
import json
import os
import sys
import time
from functools import partial

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dvst_b200  # noqa
from dvst_b200 import _lib
from dvst_b200.dino_similarity import DINOLoss, HeadScorer, preprocess_frames, score_video, score_video_device
from dvst_b200.models.timesformer import VisionTransformer
from dvst_b200.testing import synth_backbone_state_dict, synth_clips, synth_head_state_dict, synth_video
from dvst_b200.vision_transformer import DINOHead

dev = torch.device("cuda:0")
ok = True
G = torch.load(os.path.join(ROOT, "tests", "golden", "head_loss.pt"))


def build(prec, depth=12, seed=0):
    m = VisionTransformer(img_size=224, num_classes=0, depth=depth, qkv_bias=True,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8, precision=prec)
    m.load_state_dict(synth_backbone_state_dict(seed, depth=depth, num_frames=8), strict=True)
    return m.eval().to(dev)


def relv(a, b):
    return ((a.double().cpu() - b.double().cpu()).abs() / b.double().cpu().abs()).max().item()


# ---- losses on the GOLDEN features (isolates head / loss kernels from backbone error) ----
fs, ft = G["feat_s"].to(dev), G["feat_t"].to(dev)
lossR = DINOLoss(768, 0.02, 0.3).to(dev)
e = relv(lossR.pair_losses(fs, ft), G["loss_R"])
print(f"mode-R loss kernel on golden feats: max rel {e:.3e}"); ok &= e < 1e-5
for out_dim in (65536, 4096):
    for prec in ("bf16", "high"):
        head = DINOHead(768, out_dim, precision=prec)
        head.load_state_dict(synth_head_state_dict(2, out_dim=out_dim), strict=True)
        head = head.to(dev).eval()
        logits = head(fs)
        e1 = (logits[:, :8].cpu() - G[f"logits_s_{out_dim}_head8"]).abs().max().item()
        e2 = (logits.sum(1).cpu() - G[f"logits_s_{out_dim}_sum"]).abs().max().item()
        print(f"DINOHead({out_dim}) {prec}: logits max abs err {e1:.3e}, rowsum err {e2:.3e}")
        ok &= e1 < (5e-3 if prec == "bf16" else 1e-4)
        g = torch.Generator().manual_seed(7)
        center = torch.randn(1, out_dim, generator=g) * 0.05
        assert torch.equal(center[0, :8], G[f"center_seed7_{out_dim}_head8"])
        idx = torch.arange(4, dtype=torch.int32, device=dev)
        for (tt, st, cname) in [(0.04, 0.1, "zero"), (0.04, 0.1, "rand"), (0.02, 0.3, "zero"), (0.07, 0.1, "rand")]:
            sc = HeadScorer(head, tt, st, None if cname == "zero" else center.to(dev))
            got = sc.pair_losses(fs, ft, idx, idx)
            e = relv(got, G[f"loss_H{out_dim}_{tt}_{st}_{cname}"])
            # materialised-logits path through the plain loss kernel
            L2 = DINOLoss(out_dim, tt, st).to(dev)
            if cname == "rand":
                L2.center.copy_(center)
            e_m = relv(L2.pair_losses(head(fs), head(ft)), G[f"loss_H{out_dim}_{tt}_{st}_{cname}"])
            print(f"  mode-H fused loss tt={tt} ts={st} center={cname}: max rel {e:.3e} (materialised: {e_m:.3e})")
            ok &= e < 1e-3

# ---- end-to-end pairs: backbone + losses vs golden (BASELINE config 1: 4 candidate subsets) ----
xs, xt = synth_clips(200, 4, 3).to(dev), synth_clips(201, 4, 8).to(dev)
for prec in ("bf16", "high"):
    m = build(prec)
    with torch.no_grad():
        f_s, f_t = m(xs), m(xt)
    eR = relv(lossR.pair_losses(f_s, f_t), G["loss_R"])
    head = DINOHead(768, 65536, precision=prec)
    head.load_state_dict(synth_head_state_dict(2), strict=True)
    head = head.to(dev).eval()
    idx = torch.arange(4, dtype=torch.int32, device=dev)
    eH = relv(HeadScorer(head, 0.04, 0.1).pair_losses(f_s, f_t, idx, idx), G["loss_H65536_0.04_0.1_zero"])
    print(f"end-to-end {prec}: mode-R loss max rel {eR:.3e}; mode-H loss max rel {eH:.3e}")
    ok &= eH < 1e-3

# ---- whole-driver golden (reference's own loop) ----
D = json.load(open(os.path.join(ROOT, "tests", "golden", "driver_loss.json")))
for prec in ("bf16", "high"):
    m = build(prec, seed=D["seed_w"])
    for name, meta in D["videos"].items():
        g = torch.Generator().manual_seed(meta["seed"])
        raw = torch.randint(0, 256, (meta["n_unsampled"], 224, 224, 3), generator=g, dtype=torch.uint8)
        frames = preprocess_frames(raw.to(dev), 4)
        with torch.no_grad():
            got = score_video(m, frames)
        ref = D["loss"][name]
        e = max(abs(a - b) / abs(b) for a, b in zip(got, ref))
        print(f"driver {name} {prec}: F={len(ref)} max rel {e:.3e} inv(v0==v1)={got[0] == got[1]} "
              f"inv(last)={got[-1] == got[-2]}")
        ok &= len(got) == len(ref) and got[0] == got[1] and got[-1] == got[-2]

# ---- dummy path and timing ----
m = build("bf16")
with torch.no_grad():
    d = score_video(m, torch.zeros(5, 3, 112, 112, device=dev))  # wrong crop -> dummy views
print("dummy video:", len(d), "values, all equal:", len(set(d)) == 1, d[0])
ok &= len(d) == 30 and len(set(d)) == 1
for F in (30, 68, 75):
    frames = synth_video(5, F).to(dev)
    with torch.no_grad():
        score_video_device(m, frames, lossR)
        torch.cuda.synchronize()
        t0 = time.time()
        n0 = _lib.launch_count()
        for _ in range(3):
            out = score_video_device(m, frames, lossR)
        torch.cuda.synchronize()
    dt = (time.time() - t0) / 3
    print(f"score_video F={F}: {dt * 1e3:.1f} ms -> {1 / dt:.2f} videos/s, launches={(_lib.launch_count() - n0) // 3}")
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
