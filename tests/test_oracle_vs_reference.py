"""Oracle restatement vs the imported reference in fp64 (authoring container only: /root/reference does not exist on
the GPU box, where these tests skip)."""
import functools

import pytest
import torch

import dvst_b200  # noqa: F401
import dvst_oracle as O
import ref_shim
from dvst_b200.testing import synth_backbone_state_dict, synth_clips, synth_head_state_dict

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted")


@pytest.mark.parametrize("att,T,crop", [("divided_space_time", 5, 224), ("divided_space_time", 3, 96),
                                         ("joint_space_time", 2, 224), ("space_only", 2, 224)])
def test_forward_fp64(att, T, crop):
    R = ref_shim.load()
    sd = synth_backbone_state_dict(1, depth=1, num_frames=8, attention_type=att)
    m = R["VisionTransformer"](img_size=224, num_classes=0, patch_size=16, embed_dim=768, depth=1, num_heads=12,
                               mlp_ratio=4, qkv_bias=True, norm_layer=functools.partial(torch.nn.LayerNorm, eps=1e-6),
                               num_frames=8, attention_type=att)
    m.load_state_dict(sd, strict=True)
    m = m.eval().double()
    x = synth_clips(3, 2, T, crop, crop).double()
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        ref = m.forward_features(x, get_all=True)
    got = O.forward_features(sd64, x, att, get_all=True)
    assert (got - ref).abs().max().item() < 1e-12


def test_attention_map_fp64():
    R = ref_shim.load()
    sd = synth_backbone_state_dict(1, depth=2, num_frames=8)
    m = R["VisionTransformer"](img_size=224, num_classes=0, patch_size=16, embed_dim=768, depth=2, num_heads=12,
                               mlp_ratio=4, qkv_bias=True, norm_layer=functools.partial(torch.nn.LayerNorm, eps=1e-6),
                               num_frames=8, attention_type="divided_space_time")
    m.load_state_dict(sd, strict=True)
    m = m.eval().double()
    x = synth_clips(4, 1, 3, 96, 96).double()
    with torch.no_grad():
        ref = m.forward_features(x, get_attn=True)
    got = O.attention_map({k: v.double() for k, v in sd.items()}, x)
    assert got.shape == ref.shape and (got - ref).abs().max().item() < 1e-13


def test_head_and_loss_fp64():
    R = ref_shim.load()
    hsd = synth_head_state_dict(2, out_dim=1024)
    head = R["DINOHead"](768, 1024)
    head.load_state_dict(hsd, strict=True)
    head = head.double().eval()
    x = torch.randn(3, 768, dtype=torch.float64)
    with torch.no_grad():
        ref = head(x)
    got = O.dino_head({k: v.double() for k, v in hsd.items()}, x)
    assert (got - ref).abs().max().item() < 1e-12
    L = R["DINOLoss"](out_dim=1024, teacher_temp=0.04, student_temp=0.1).double()
    with torch.no_grad():
        lref = torch.stack([L(ref[k], ref[(k + 1) % 3]) for k in range(3)])
    lgot = O.dino_pair_loss(got, got.roll(-1, 0), torch.zeros(1, 1024, dtype=torch.float64), 0.04, 0.1)
    assert (lgot - lref).abs().max().item() < 1e-10


def test_knn_classifier_live():
    """oracle knn_classifier vs the reference function compiled from its own source (eval_knn.py:138-178)."""
    from dvst_b200.testing import synth_knn_features
    knn = ref_shim.load_knn_classifier()
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for seed, k, nc in ((11, 7, 9), (12, 30, 25)):
            trf, trl, tef, tel = synth_knn_features(seed, 500, 200, 48, nc, 3.0)
            ref = knn(trf, trl, tef, tel, k, 0.07, nc)
            got = O.knn_classifier(trf, trl, tef, tel, k, 0.07, nc)
            assert abs(got[0] - ref[0]) < 1e-9 and abs(got[1] - ref[1]) < 1e-9
    finally:
        torch.Tensor.cuda = orig
