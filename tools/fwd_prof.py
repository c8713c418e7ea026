"""One cfg2-shaped forward (target of ncu captures / launch lists)."""
import os, sys, torch
from functools import partial
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
from dvst_b200.models.timesformer import VisionTransformer
from dvst_b200.testing import synth_backbone_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 12
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda:0")
if len(sys.argv) > 5:
    _lib.lib().dvst_set_chunk_rows(int(sys.argv[5]))
prec = os.environ.get("DVST_PRECISION", "bf16")   # "high" = 2-term bf16 split
m = VisionTransformer(img_size=224, num_classes=0, depth=depth, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8,
                      precision=prec)
m.load_state_dict(synth_backbone_state_dict(0, depth=depth, num_frames=8), strict=True)
m = m.eval().to(dev)
x = torch.randn(B, 3, T, 224, 224, device=dev)
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    _lib.profile_enable(True); _lib.profile_collect()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m(x)
    e1.record(); torch.cuda.synchronize()
print("forward ms", e0.elapsed_time(e1) / reps, {k: (round(v[0] / reps, 3), v[2] // reps) for k, v in _lib.profile_collect().items()})
