"""Context number for DESIGN.md: the oracle (plain torch restatement of the reference's forward) run with torch EAGER on
the same B200 - fp32 (TF32 off, the reference's arithmetic), TF32 and autocast(bf16) - next to dvst_b200 on the cfg2 step
shape.  Test infrastructure (imports oracle/): not part of the product path or of bench.py's JSON line.
    python tools/eager_baseline.py [clips]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dvst_b200  # noqa: F401
import dvst_oracle as O
from dvst_b200.testing import synth_backbone_state_dict, synth_clips
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
sd = {k: v.to(dev) for k, v in synth_backbone_state_dict(0, depth=12, num_frames=8).items()}
x = synth_clips(1, B, 16).to(dev)


def run(label, ctx):
    with torch.no_grad(), ctx:
        O.forward_features(sd, x); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            f = O.forward_features(sd, x)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{label:28s} {B / dt:8.1f} clips/s ({dt * 1e3:.1f} ms for {B} clips x 16 frames)", flush=True)
    return f.float()


import contextlib
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
f32 = run("torch eager fp32 (no TF32)", contextlib.nullcontext())
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
run("torch eager TF32", contextlib.nullcontext())
fb = run("torch eager autocast(bf16)", torch.autocast("cuda", dtype=torch.bfloat16))
print("autocast(bf16) feature rel-L2 vs fp32:", ((fb - f32).norm() / f32.norm()).item())
from functools import partial
from dvst_b200.models.timesformer import VisionTransformer
m = VisionTransformer(img_size=224, num_classes=0, depth=12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
m.load_state_dict({k: v.cpu() for k, v in sd.items()}, strict=True); m = m.eval().to(dev)
with torch.no_grad():
    m(x); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fo = m(x)
    torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"{'dvst_b200 (bf16)':28s} {B / dt:8.1f} clips/s; feature rel-L2 vs eager fp32: {((fo - f32).norm() / f32.norm()).item():.2e}")
