"""The oracle (oracle/dvst_oracle.py) against the golden fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import json
import os

import pytest
import torch

import dvst_b200  # noqa: F401
import dvst_oracle as O
from dvst_b200.testing import synth_backbone_state_dict, synth_clips, synth_head_state_dict


@pytest.fixture(scope="module")
def backbone_gold(golden_dir):
    g = torch.load(os.path.join(golden_dir, "backbone.pt"))
    g.update(torch.load(os.path.join(golden_dir, "backbone_r2.pt")))
    return g


@pytest.mark.parametrize("name", ["cfg1_full_T8", "d2_T5", "d2_T30", "d2_T16", "d2_T4_crop96", "d2_T1", "d2_joint_T4",
                                  "d2_space_T4", "cfg2_full_T16", "cfg5_d2_joint_T32"])
def test_backbone_matches_reference(backbone_gold, name):
    c = backbone_gold[name]
    sd = synth_backbone_state_dict(c["seed_w"], depth=c["depth"], num_frames=c["num_frames"],
                                   attention_type=c["attention_type"])
    x = synth_clips(c["seed_x"], c["B"], c["T"], c["crop"], c["crop"])
    with torch.no_grad():
        f = O.forward_features(sd, x, c["attention_type"])
    # both sides are fp32 CPU: differences are summation-order noise
    assert (f - c["feat"]).abs().max().item() < 2e-4
    assert ((f - c["feat"]).norm() / c["feat"].norm()).item() < 2e-5


def test_attention_map_matches_reference(golden_dir):
    """forward_features(get_attn=True): the last block's spatial attention map (models/timesformer.py:325-331)."""
    G = torch.load(os.path.join(golden_dir, "attn_map.pt"))
    for name, c in G.items():
        sd = synth_backbone_state_dict(c["seed_w"], depth=c["depth"], num_frames=c["num_frames"])
        x = synth_clips(c["seed_x"], c["B"], c["T"], c["crop"], c["crop"])
        with torch.no_grad():
            a = O.attention_map(sd, x)
        if "rows" in c:
            a = a[:, :, c["rows"], :]
        assert a.shape == c["attn"].shape
        assert (a - c["attn"]).abs().max().item() < 1e-5
        assert (a.sum(-1) - 1).abs().max().item() < 1e-5


def test_head_and_losses_match_reference(golden_dir):
    G = torch.load(os.path.join(golden_dir, "head_loss.pt"))
    fs, ft = G["feat_s"], G["feat_t"]
    lr = O.dino_pair_loss(fs, ft, torch.zeros(1, 768), 0.02, 0.3)
    assert torch.allclose(lr, G["loss_R"], rtol=1e-5, atol=1e-6)
    for out_dim in (65536, 4096):
        hsd = synth_head_state_dict(2, out_dim=out_dim)
        ls, lt = O.dino_head(hsd, fs), O.dino_head(hsd, ft)
        assert torch.allclose(ls[:, :8], G[f"logits_s_{out_dim}_head8"], rtol=1e-4, atol=1e-6)
        assert torch.allclose(lt.sum(0)[:8], G[f"logits_t_{out_dim}_colsum_head8"], rtol=1e-4, atol=1e-5)
        g = torch.Generator().manual_seed(7)
        center = torch.randn(1, out_dim, generator=g) * 0.05
        for (tt, st, cname) in [(0.04, 0.1, "zero"), (0.04, 0.1, "rand"), (0.02, 0.3, "zero"), (0.07, 0.1, "rand")]:
            c = torch.zeros(1, out_dim) if cname == "zero" else center
            got = O.dino_pair_loss(ls, lt, c, tt, st)
            assert torch.allclose(got, G[f"loss_H{out_dim}_{tt}_{st}_{cname}"], rtol=2e-5, atol=1e-5)


def test_window_rule_matches_reference(golden_dir):
    W = json.load(open(os.path.join(golden_dir, "windows.json")))
    for F, c in W.items():
        G, starts = O.window_starts(int(F))
        assert G == c["G"]
        assert [list(s) for s in starts] == [list(s) for s in c["starts"]]


def test_selection_matches_reference(golden_dir):
    S = json.load(open(os.path.join(golden_dir, "selection.json")))
    assert len(S) == 144
    for name, c in S.items():
        for N in (8, 16):
            assert O.select_frames(c["loss"], c["n_unsampled"], 4, N) == c[f"idx{N}"], name


def test_driver_loop_matches_reference(golden_dir):
    """The reference's own dino_similarity() loop (loader -> views -> model -> DINOLoss -> JSON) on two tiny videos."""
    preprocess_frames = O.preprocess_frames
    D = json.load(open(os.path.join(golden_dir, "driver_loss.json")))
    sd = synth_backbone_state_dict(D["seed_w"], depth=12, num_frames=8)
    for name, meta in D["videos"].items():
        g = torch.Generator().manual_seed(meta["seed"])
        raw = torch.randint(0, 256, (meta["n_unsampled"], 224, 224, 3), generator=g, dtype=torch.uint8)
        frames = preprocess_frames(raw, 4)
        with torch.no_grad():
            got = O.score_video(sd, frames)
        ref = D["loss"][name]
        assert len(got) == len(ref)
        assert max(abs(a - b) / abs(b) for a, b in zip(got, ref)) < 2e-4
        assert got[0] == got[1] and got[-1] == got[-2]  # structural invariants of the shipped JSON


def test_nearest_index_matches_aten():
    import torch.nn.functional as F
    for src in (8, 14, 16, 32):
        for dst in list(range(1, 70)):
            ref = F.interpolate(torch.arange(src, dtype=torch.float32)[None, None], size=dst, mode="nearest")[0, 0]
            assert O.nearest_index(dst, src).tolist() == ref.long().tolist(), (src, dst)


def test_center_update_and_train_loss_shapes():
    torch.manual_seed(0)
    out_dim, B = 512, 3
    s, t = torch.randn(4 * B, out_dim), torch.randn(2 * B, out_dim)
    c = torch.randn(1, out_dim) * 0.1
    l = O.dino_train_loss(s, t, c, 0.04, 0.1, ncrops=4)
    assert l.ndim == 0 and torch.isfinite(l)
    c2 = O.center_update(c, [t[:B], t[B:]])
    assert torch.allclose(c2, c * 0.9 + t.sum(0, keepdim=True) / (B * 2) * 0.1, atol=1e-6)


def test_knn_classifier_matches_reference(golden_dir):
    """oracle knn_classifier against the reference's own knn_classifier outputs (eval_knn.py:138-178)."""
    from dvst_b200.testing import synth_knn_features
    cases = json.load(open(os.path.join(golden_dir, "knn.json")))
    assert len(cases) >= 6
    for c in cases:
        trf, trl, tef, tel = synth_knn_features(c["seed"], c["n_train"], c["n_test"], c["dim"], c["num_classes"],
                                                c["noise"])
        top1, top5, rank, idx = O.knn_classifier(trf, trl, tef, tel, c["k"], c["T"], c["num_classes"])
        assert abs(top1 - c["top1"]) < 1e-9 and abs(top5 - c["top5"]) < 1e-9, c
        assert idx.shape == (c["n_test"], c["k"]) and rank.shape == (c["n_test"],)
