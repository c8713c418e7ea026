import os, sys, torch
from functools import partial
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
from dvst_b200.models.timesformer import VisionTransformer
from dvst_b200.testing import synth_backbone_state_dict
dev = torch.device("cuda:0")
for att, B, T in (("joint_space_time", 8, 32), ("joint_space_time", 16, 8), ("space_only", 64, 16)):
    m = VisionTransformer(img_size=224, num_classes=0, depth=12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8, attention_type=att)
    m.load_state_dict(synth_backbone_state_dict(0, depth=12, num_frames=8, attention_type=att), strict=True)
    m = m.eval().to(dev)
    x = torch.randn(B, 3, T, 224, 224, device=dev)
    with torch.no_grad():
        m(x); torch.cuda.synchronize()
        _lib.profile_enable(True); _lib.profile_collect()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2): m(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    prof = {k: round(v[0] / 2, 2) for k, v in _lib.profile_collect().items()}
    _lib.profile_enable(False)
    print(att, "B", B, "T", T, "forward ms", round(ms, 2), "clips/s", round(B / ms * 1e3, 1), prof, flush=True)
    del m
