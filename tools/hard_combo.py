"""Feature error vs the fp32 oracle for every combination of {deferred LayerNorm on / off} x {attention: one-pass tcgen05,
two-pass tcgen05, mma.sync} on the easy and the hard synthetic weight sets (dvst_b200.testing).  Test infrastructure."""
import sys, torch
from functools import partial
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import dvst_b200, dvst_oracle as O
from dvst_b200 import _lib
from dvst_b200.models.timesformer import VisionTransformer
from dvst_b200.testing import synth_backbone_state_dict, synth_clips
L=_lib.lib(); dev=torch.device("cuda:0")
for hard in (True, False):
  for seed in (5, 6):
    sd = synth_backbone_state_dict(seed, depth=4, num_frames=8, hard=hard)
    x = synth_clips(55+seed, 2, 4)
    with torch.no_grad(): ref = O.forward_features(sd, x)
    m = VisionTransformer(img_size=224, num_classes=0, depth=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
    m.load_state_dict(sd, strict=True); m = m.eval().to(dev)
    for ln in (1,0):
        for att in (1,2,0):
            L.dvst_set_ln_fusion(ln); L.dvst_set_attention_impl(att)
            with torch.no_grad(): f = m(x.to(dev))
            print("hard" if hard else "easy", seed, "ln_fusion", ln, "attn_impl", att, "rel", ((f.cpu()-ref).norm()/ref.norm()).item())
