"""End-to-end parity of the CUDA path on a B200: against the golden fixtures produced by the unmodified reference,
against the CPU oracle on fresh inputs, and - at BASELINE.json's full sizes - through size-independent properties."""
import json
import os
from functools import partial

import pytest
import torch

import dvst_b200  # noqa: F401
import dvst_oracle as O
from dvst_b200 import _lib
from dvst_b200.dino_similarity import (DINOLoss, HeadScorer, preprocess_frames, score_video, score_video_device,
                                       score_videos_device)
from dvst_b200.frame_selection import select_frames_adaptive
from dvst_b200.models.timesformer import VisionTransformer
from dvst_b200.testing import synth_backbone_state_dict, synth_clips, synth_head_state_dict, synth_video
from dvst_b200.train_loss import DINOLoss as TrainDINOLoss
from dvst_b200.vision_transformer import DINOHead

pytestmark = pytest.mark.gpu

FEAT_TOL = 2e-2      # north star: features within 2e-2 relative (bf16)
LOSS_TOL = 1e-3      # north star: per-subset DINO loss within 1e-3 relative (mode H)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


_models = {}


def build(dev, seed, depth, att="divided_space_time", prec="bf16", T0=8):
    key = (seed, depth, att, prec, T0)
    if key not in _models:
        m = VisionTransformer(img_size=224, num_classes=0, depth=depth, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=T0, attention_type=att,
                              precision=prec)
        m.load_state_dict(synth_backbone_state_dict(seed, depth=depth, num_frames=T0, attention_type=att), strict=True)
        _models[key] = m.eval().to(dev)
    return _models[key]


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()


def _relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs() / b.abs()).max().item()


GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "backbone.pt"))
# round 2: the benchmark's own shapes at full depth, produced by the unmodified reference (oracle/make_golden.py):
# cfg2 (depth 12, T = 16), cfg3's teacher window (depth 12, T = 30), cfg5 (depth 12 divided T = 32; joint T = 32 with
# 6 273 tokens = 49 key tiles of the flash-style kernel)
GOLD.update(torch.load(os.path.join(os.path.dirname(__file__), "golden", "backbone_r2.pt")))


@pytest.mark.parametrize("name", sorted(GOLD))
@pytest.mark.parametrize("prec", ["bf16", "high"])
def test_backbone_vs_reference_golden(dev, name, prec):
    c = GOLD[name]
    m = build(dev, c["seed_w"], c["depth"], c["attention_type"], prec, c["num_frames"])
    x = synth_clips(c["seed_x"], c["B"], c["T"], c["crop"], c["crop"]).to(dev)
    with torch.no_grad():
        f = m(x)
    assert tuple(f.shape) == (c["B"], 768) and f.dtype == torch.float32
    assert _rel(f, c["feat"]) < (FEAT_TOL if prec == "bf16" else 1e-4)


def test_get_all_token_order_vs_oracle(dev):
    sd = synth_backbone_state_dict(1, depth=2, num_frames=8)
    m = build(dev, 1, 2)
    x = synth_clips(9, 2, 3)
    ref = O.forward_features(sd, x, get_all=True)
    with torch.no_grad():
        got = m.forward_features(x.to(dev), get_all=True)
        got2 = m.get_intermediate_layers(x.to(dev))[0]
    assert tuple(got.shape) == (2, 1 + 196 * 3, 768)
    assert _rel(got, ref) < FEAT_TOL and torch.equal(got, got2)
    assert _rel(got[:, 0], ref[:, 0]) < FEAT_TOL and _rel(got[:, 5], ref[:, 5]) < FEAT_TOL


@pytest.mark.parametrize("att", ["divided_space_time", "joint_space_time", "space_only"])
@pytest.mark.parametrize("prec", ["bf16", "high"])
def test_cls_only_last_block_matches_full_forward(dev, att, prec):
    """forward() carries only the CLS rows through the last block's spatial / MLP half (the patch tokens of that block
    feed nothing, timesformer.py:342-345); the result must equal the CLS row of the full get_all forward, where every
    token is computed: fp32 CUDA-core CLS attention vs tensor-core attention differ by bf16 rounding of P only."""
    m = build(dev, 3, 2, att, prec)
    x = synth_clips(21, 2, 4).to(dev)
    with torch.no_grad():
        cls_only = m(x)
        full = m.forward_features(x, get_all=True)[:, 0]
    assert _rel(cls_only, full) < (3e-3 if prec == "bf16" else 2e-5)


def test_hard_weights_stay_within_the_feature_budget(dev):
    """Weights shaped like a trained checkpoint's worst habits (sharp attention, LayerNorm gains up to 3, a massive-
    activation channel, token rows whose mean is several times their spread): bf16 features stay within the 2e-2 budget
    of the fp32 oracle with the deferred LayerNorm / one-pass softmax (default) and with the stand-alone kernels, and
    precision="high" stays tight."""
    from dvst_b200 import _lib
    L = _lib.lib()
    sd = synth_backbone_state_dict(5, depth=4, num_frames=8, hard=True)
    x = synth_clips(55, 2, 4)
    with torch.no_grad():
        ref = O.forward_features(sd, x)
    row = O.embed_tokens(sd, x)[0].reshape(-1, 768)
    assert float((row.mean(1).abs() / row.std(1)).median()) > 0.5   # the offsets really are large next to the spread
    # yardstick: the error ANY implementation with bf16 weights has on this ill-conditioned set (weights rounded to
    # bf16, everything else fp32 on the CPU)
    with torch.no_grad():
        yard = _rel(O.forward_features({k: (v.bfloat16().float() if v.dim() >= 2 else v) for k, v in sd.items()}, x), ref)
    assert 5e-3 < yard < 1.5e-2
    for prec, tol in (("bf16", max(FEAT_TOL, 2.5 * yard)), ("high", 5e-4)):
        m = VisionTransformer(img_size=224, num_classes=0, depth=4, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8, precision=prec)
        m.load_state_dict(sd, strict=True)
        m = m.eval().to(dev)
        try:
            with torch.no_grad():
                f1 = m(x.to(dev))
                L.dvst_set_ln_fusion(0)
                L.dvst_set_attention_impl(2)
                f0 = m(x.to(dev))
        finally:
            L.dvst_set_ln_fusion(1)
            L.dvst_set_attention_impl(1)
        print(prec, "fused", _rel(f1, ref), "stand-alone", _rel(f0, ref), "weight-rounding yardstick", yard)
        assert _rel(f1, ref) < tol and _rel(f0, ref) < tol, (prec, _rel(f1, ref), _rel(f0, ref), yard)


def test_deferred_layernorm_matches_standalone_layernorm(dev):
    """bf16 precision runs the deferred-LayerNorm path by default (residual GEMMs emit bf16(x) + row statistics, the
    next GEMM applies the normalisation in its epilogue); forcing the stand-alone LayerNorm kernel must give the same
    features up to bf16 rounding, for every attention type and for the get_all / CLS-only outputs."""
    from dvst_b200 import _lib
    L = _lib.lib()
    for att in ("divided_space_time", "joint_space_time", "space_only"):
        m = build(dev, 4, 2, att, "bf16")
        x = synth_clips(41, 3, 4).to(dev)
        try:
            with torch.no_grad():
                L.dvst_set_ln_fusion(0)
                n0 = _lib.launch_count()
                f0, a0 = m(x), m.forward_features(x, get_all=True)
                n_plain = _lib.launch_count() - n0
                L.dvst_set_ln_fusion(1)
                n0 = _lib.launch_count()
                f1, a1 = m(x), m.forward_features(x, get_all=True)
                n_fused = _lib.launch_count() - n0
        finally:
            L.dvst_set_ln_fusion(1)
        assert n_fused < n_plain                       # the LayerNorm launches are gone
        assert _rel(f1, f0) < 1.5e-2 and _rel(a1, a0) < 1.5e-2   # two bf16 paths, each ~6e-3 from fp32


@pytest.mark.parametrize("B,T", [(3, 4), (10, 16)])   # (10, 16): 31 360 rows - the 2-CTA 256-row tiles
def test_deferred_layernorm_in_high_precision(dev, B, T):
    """precision="high" runs the same deferred-LayerNorm form with every operand a bf16 [hi | lo] pair: the residual GEMMs
    leave xb = [bf16(x - m) | bf16 of the remainder], the consumers fold the normalisation around their three-product
    accumulator.  Against the stand-alone LayerNorm kernel the features agree to the precision of the split (1e-5), far inside
    the 1e-4 budget of the reference goldens; the LayerNorm launches are gone."""
    from dvst_b200 import _lib
    L = _lib.lib()
    for att in ("divided_space_time", "joint_space_time", "space_only") if B == 3 else ("divided_space_time",):
        m = build(dev, 4, 2, att, "high")
        x = synth_clips(41, B, T).to(dev)
        try:
            with torch.no_grad():
                L.dvst_set_ln_fusion(0)
                n0 = _lib.launch_count()
                f0, a0 = m(x), m.forward_features(x, get_all=True)
                n_plain = _lib.launch_count() - n0
                L.dvst_set_ln_fusion(1)
                n0 = _lib.launch_count()
                f1, a1 = m(x), m.forward_features(x, get_all=True)
                n_fused = _lib.launch_count() - n0
        finally:
            L.dvst_set_ln_fusion(1)
        assert n_fused < n_plain
        assert _rel(f1, f0) < 5e-5 and _rel(a1, a0) < 5e-5, (att, _rel(f1, f0), _rel(a1, a0))


@pytest.mark.parametrize("T,crop,same_bits", [(16, 224, True), (8, 96, True), (32, 96, True), (4, 96, True),
                                              (3, 224, False), (30, 96, False), (8, 112, False), (5, 96, False)])
def test_fused_temporal_half_matches_two_kernel_form(dev, T, crop, same_bits):
    """The temporal half of a divided block (temporal_norm1 -> qkv -> attention) runs as ONE kernel for T <= 64 (qkv GEMM
    whose epilogue does the temporal attention out of shared memory, csrc/gemm_tattn_sm100.cu): one launch fewer per block.
    Power-of-two T on 16-row-aligned clips: the SAME BITS as the two-kernel form (same arithmetic in the same order).
    Other shapes (the 4_3_30 windows T = 3 / 30, 8 x 7 x 7 tokens): rows are tiled clip by clip, the result agrees to bf16
    rounding.  In every case the features do not depend on the batch composition."""
    from dvst_b200 import _lib
    L = _lib.lib()
    m = build(dev, 5, 3, "divided_space_time", "bf16")
    x = synth_clips(43, 3, T, crop, crop).to(dev)
    try:
        with torch.no_grad():
            L.dvst_set_temporal_fusion(0)
            n0 = _lib.launch_count()
            f0, a0 = m(x), m.forward_features(x, get_all=True)
            n_two = _lib.launch_count() - n0
            L.dvst_set_temporal_fusion(1)
            n0 = _lib.launch_count()
            f1, a1 = m(x), m.forward_features(x, get_all=True)
            n_one = _lib.launch_count() - n0
            single = torch.cat([m(x[i:i + 1]) for i in range(3)])
            rev = m(x.flip(0)).flip(0)
    finally:
        L.dvst_set_temporal_fusion(1)
    assert n_two - n_one == 2 * 3                      # one launch fewer per block and forward
    if same_bits:
        assert torch.equal(f0, f1) and torch.equal(a0, a1)
    else:
        assert _rel(f1, f0) < 6e-3 and _rel(a1, a0) < 6e-3
    assert torch.equal(single, f1) and torch.equal(rev, f1)   # batch-composition independence


@pytest.mark.parametrize("prec", ["bf16", "high"])
def test_attention_map_vs_reference_golden(dev, prec):
    """forward_features(x, get_attn=True): last block's spatial attention map, against the reference's own output."""
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "attn_map.pt"))
    for name, c in G.items():
        m = build(dev, c["seed_w"], c["depth"], "divided_space_time", prec, c["num_frames"])
        x = synth_clips(c["seed_x"], c["B"], c["T"], c["crop"], c["crop"]).to(dev)
        with torch.no_grad():
            a = m.forward_features(x, get_attn=True)
        N = (c["crop"] // 16) ** 2
        assert tuple(a.shape) == (c["B"] * c["T"], 12, N + 1, N + 1) and a.dtype == torch.float32
        assert (a.sum(-1) - 1).abs().max().item() < 1e-5
        if "rows" in c:
            a = a[:, :, c["rows"], :]
        assert _rel(a, c["attn"]) < (FEAT_TOL if prec == "bf16" else 1e-4)
    with pytest.raises(ValueError):
        build(dev, 2, 2, "joint_space_time")(x)  # sanity: model builds; get_attn on it must raise
        build(dev, 2, 2, "joint_space_time").forward_features(x, get_attn=True)


def test_reference_smoke_shape(dev):
    """models/timesformer.py:639-643: ones [2,3,8,224,224] -> [2,768]."""
    m = build(dev, 1, 2)
    with torch.no_grad():
        out = m(torch.ones(2, 3, 8, 224, 224, device=dev))
    assert tuple(out.shape) == (2, 768) and torch.isfinite(out).all()
    assert torch.equal(out[0], out[1])  # identical clips -> bit-identical features


def test_losses_vs_reference_golden(dev):
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "head_loss.pt"))
    fs, ft = G["feat_s"].to(dev), G["feat_t"].to(dev)
    assert _relmax(DINOLoss(768, 0.02, 0.3).to(dev).pair_losses(fs, ft), G["loss_R"]) < 1e-5
    single = DINOLoss(768, 0.02, 0.3).to(dev)(fs[1], ft[1])  # reference call form: two 768-vectors -> 0-dim tensor
    assert single.ndim == 0 and abs(float(single) - float(G["loss_R"][1])) < 1e-5 * float(G["loss_R"][1])
    for out_dim in (65536, 4096):
        head = DINOHead(768, out_dim)
        head.load_state_dict(synth_head_state_dict(2, out_dim=out_dim), strict=True)
        head = head.to(dev).eval()
        logits = head(fs)
        assert tuple(logits.shape) == (4, out_dim)
        assert (logits[:, :8].cpu() - G[f"logits_s_{out_dim}_head8"]).abs().max().item() < 3e-3
        g = torch.Generator().manual_seed(7)
        center = torch.randn(1, out_dim, generator=g) * 0.05
        idx = torch.arange(4, dtype=torch.int32, device=dev)
        for (tt, st, cname) in [(0.04, 0.1, "zero"), (0.04, 0.1, "rand"), (0.02, 0.3, "zero"), (0.07, 0.1, "rand")]:
            sc = HeadScorer(head, tt, st, None if cname == "zero" else center.to(dev))
            assert _relmax(sc.pair_losses(fs, ft, idx, idx), G[f"loss_H{out_dim}_{tt}_{st}_{cname}"]) < LOSS_TOL


def test_config1_end_to_end(dev):
    """BASELINE config 1: ViT-B/16 8x224^2 batch 2 + DINO loss over 4 candidate subsets (student T=3, teacher T=8)."""
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "head_loss.pt"))
    xs, xt = synth_clips(200, 4, 3).to(dev), synth_clips(201, 4, 8).to(dev)
    idx = torch.arange(4, dtype=torch.int32, device=dev)
    head = DINOHead(768, 65536)
    head.load_state_dict(synth_head_state_dict(2), strict=True)
    head = head.to(dev).eval()
    for prec, tol_r in (("bf16", 5e-2), ("high", LOSS_TOL)):
        m = build(dev, 0, 12, prec=prec)
        with torch.no_grad():
            fs, ft = m(xs), m(xt)
        assert _rel(fs, G["feat_s"]) < FEAT_TOL and _rel(ft, G["feat_t"]) < FEAT_TOL
        # mode H (north-star contract): loss within 1e-3
        assert _relmax(HeadScorer(head, 0.04, 0.1).pair_losses(fs, ft, idx, idx), G["loss_H65536_0.04_0.1_zero"]) < LOSS_TOL
        # mode R (tau_t = 0.02 amplifies feature error 50x; SURVEY finding 3): bf16 operands cannot meet 1e-3 (stated,
        # measured ~2e-2); precision="high" (2-term bf16 split in every GEMM and both attentions) does
        assert _relmax(DINOLoss(768, 0.02, 0.3).to(dev).pair_losses(fs, ft), G["loss_R"]) < tol_r


def test_driver_golden_and_invariants(dev):
    D = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "driver_loss.json")))
    m = build(dev, D["seed_w"], 12, prec="high")
    for name, meta in D["videos"].items():
        g = torch.Generator().manual_seed(meta["seed"])
        raw = torch.randint(0, 256, (meta["n_unsampled"], 224, 224, 3), generator=g, dtype=torch.uint8)
        with torch.no_grad():
            got = score_video(m, preprocess_frames(raw.to(dev), 4))
        ref = D["loss"][name]
        assert len(got) == len(ref)
        assert got[0] == got[1] and got[-1] == got[-2]
        assert _relmax(got, ref) < LOSS_TOL  # reference-exact mode R through the reference's own driver loop


def test_driver_function_writes_reference_json(dev, tmp_path):
    """dvst_b200.dino_similarity.dino_similarity (the drop-in driver, batched) against the JSON the reference's own
    driver wrote for the same two videos (tests/golden/driver_loss.json): same keys, lengths, losses within 1e-3."""
    from dvst_b200.dino_similarity import dino_similarity
    D = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "driver_loss.json")))
    m = build(dev, D["seed_w"], 12, prec="high")
    data = []
    for name, meta in D["videos"].items():
        g = torch.Generator().manual_seed(meta["seed"])
        raw = torch.randint(0, 256, (meta["n_unsampled"], 224, 224, 3), generator=g, dtype=torch.uint8)
        data.append((preprocess_frames(raw.to(dev), 4), f"/some/dir/{name}.avi"))
    out = str(tmp_path / "loss.json")
    for vpb in (1, 8):
        if os.path.exists(out):
            os.remove(out)
        res = dino_similarity(None, 3, 30, 4, out, dataset=data, model=m, videos_per_batch=vpb)
        written = json.load(open(out))
        assert list(written.keys()) == list(D["loss"].keys()) and written == res
        for name, ref in D["loss"].items():
            assert len(written[name]) == len(ref) and _relmax(written[name], ref) < LOSS_TOL


def test_driver_with_the_reference_call_form(dev, tmp_path, monkeypatch):
    """dino_similarity(args, 3, 30, 4, file_path) - the reference's exact five arguments (dino_similarity.py:16,186-189),
    `args` being its argparse namespace (cfg_file / opts / pretrained_weights / num_workers / batch_size_per_gpu).  Decode
    (torchvision.io.read_video) and torch.load are mocked with the same synthetic data the reference's own driver was run
    on for tests/golden/driver_loss.json (oracle/make_golden.py:gen_driver): same keys, lengths, losses within 1e-3."""
    import types
    import torchvision.io as io
    import dvst_b200.dino_similarity as ds
    D = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "driver_loss.json")))
    raw = {}
    for name, meta in D["videos"].items():
        g = torch.Generator().manual_seed(meta["seed"])
        raw[name] = torch.randint(0, 256, (meta["n_unsampled"], 224, 224, 3), generator=g, dtype=torch.uint8)
    (tmp_path / "test.csv").write_text("".join(f"{k}.avi 0\n" for k in D["videos"]))
    yml = tmp_path / "cfg.yaml"
    yml.write_text("DATA:\n  NUM_FRAMES: 8\n  TRAIN_CROP_SIZE: 224\nTIMESFORMER:\n  ATTENTION_TYPE: 'divided_space_time'\n"
                   "TEST:\n  NUM_ENSEMBLE_VIEWS: 1\nMODEL:\n  NUM_CLASSES: 400\n")
    args = types.SimpleNamespace(cfg_file=str(yml), pretrained_weights="synthetic.pth", num_workers=0, batch_size_per_gpu=8,
                                 opts=["DATA.PATH_TO_DATA_DIR", str(tmp_path), "DATA.PATH_PREFIX", str(tmp_path)])
    sd = synth_backbone_state_dict(D["seed_w"], depth=12, num_frames=8)
    monkeypatch.setattr(io, "read_video", lambda path, pts_unit="sec": (
        raw[os.path.splitext(os.path.basename(path))[0]], None, None), raising=False)
    monkeypatch.setattr(ds.torch, "load", lambda *a, **k: {"backbone." + k_: v for k_, v in sd.items()})
    out = str(tmp_path / "loss.json")
    with torch.cuda.device(dev):
        res = ds.dino_similarity(args, 3, 30, 4, out)
    written = json.load(open(out))
    assert list(written.keys()) == list(D["loss"].keys()) and written == res
    for name, ref in D["loss"].items():
        assert len(written[name]) == len(ref) and _relmax(written[name], ref) < LOSS_TOL


def test_dummy_video_path(dev):
    """Wrong crop -> the reference's dummy views (30 equal losses of the all-zero (T=3, T=60) pair)."""
    m = build(dev, 1, 2)
    with torch.no_grad():
        b = score_video(m, torch.randn(5, 3, 112, 112, device=dev))        # wrong crop
    assert len(b) == 30 and len(set(b)) == 1
    sd = synth_backbone_state_dict(1, depth=2, num_frames=8)
    ref = O.score_video(sd, torch.zeros(5, 3, 112, 112))
    assert len(ref) == 30 and abs(b[0] - ref[0]) / abs(ref[0]) < 5e-2


def test_videos_shorter_than_the_local_window(dev):
    """F = 1, 2 < local_clip_size = 3 are NOT dummies in the reference: its loader builds broadcast views (golden
    windows_short.json, produced by the reference's get_views_of_video_same_size + the driver's [:, :, :3] slice):
    student [f1, f1] vs teacher [f0, f1] for F = 2, [f0] vs [f0] for F = 1; F equal losses."""
    from dvst_b200.dino_similarity import short_video_clips
    W = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "windows_short.json")))
    for F in (1, 2):
        si, ti = short_video_clips(F, 3)
        assert all(p["student"] == si and p["teacher"] == ti for p in W[str(F)]) and len(W[str(F)]) == F
        assert O.short_video_clips(F, 3) == (si, ti)
    m = build(dev, 1, 2, prec="high")
    sd = synth_backbone_state_dict(1, depth=2, num_frames=8)
    for F in (1, 2):
        frames = synth_video(60 + F, F, smooth=0.5)
        ref = O.score_video(sd, frames)
        with torch.no_grad():
            got = score_video(m, frames.to(dev))
            s = m(frames[short_video_clips(F)[0]].permute(1, 0, 2, 3)[None].contiguous().to(dev))
            t = m(frames.permute(1, 0, 2, 3)[None].contiguous().to(dev))
            explicit = DINOLoss(768, 0.02, 0.3).to(dev).pair_losses(s, t).item()
        assert len(got) == F == len(ref) and len(set(got)) == 1 and got[0] == explicit
        assert _relmax(got, ref) < LOSS_TOL
    with pytest.raises(ValueError):
        short_video_clips(3, 5)   # 2 frames into 3: the reference's broadcast fails too


def test_scoring_vs_oracle_and_selected_indices(dev):
    """Whole-video scoring in mode H against the oracle on fresh synthetic videos; the frame indices selected from the
    two loss curves must be identical (north star: >= 99 % of videos)."""
    sd = synth_backbone_state_dict(1, depth=2, num_frames=8)
    hsd = synth_head_state_dict(2, out_dim=4096)
    m = build(dev, 1, 2)
    head = DINOHead(768, 4096)
    head.load_state_dict(hsd, strict=True)
    head = head.to(dev).eval()
    import numpy as np
    rng = np.random.default_rng(0)

    def robust(ref, F, N):
        """the reference's own selection does not change under loss perturbations of the allowed size (LOSS_TOL): only
        then is 'identical indices' a well-posed requirement (a quantile that falls between two frames within 1e-3 flips
        with any rounding, in the reference too)"""
        base = select_frames_adaptive(ref, 4 * F, 4, N)
        return all(select_frames_adaptive([v * (1 + LOSS_TOL * rng.uniform(-1, 1)) for v in ref], 4 * F, 4, N) == base
                   for _ in range(40))

    same, total, fragile = 0, 0, 0
    for seed, F in [(1, 5), (2, 9), (3, 12), (4, 7), (5, 3), (6, 16), (7, 6), (8, 11), (9, 14), (10, 8)]:
        frames = synth_video(seed, F, smooth=0.5)
        ref = O.score_video(sd, frames, local_size=3, global_size=6, teacher_temp=0.04, student_temp=0.1, head=hsd)
        with torch.no_grad():
            got = score_video(m, frames.to(dev), local_clip_size=3, global_clip_size=6,
                              scorer=HeadScorer(head, 0.04, 0.1))
        assert len(got) == F and _relmax(got, ref) < LOSS_TOL
        for N in (2, 4, 8):
            if not robust(ref, F, N):
                fragile += 1
                continue
            total += 1
            same += select_frames_adaptive(got, 4 * F, 4, N) == select_frames_adaptive(ref, 4 * F, 4, N)
    assert total >= 24 and fragile <= 6 and same == total, (same, total, fragile)


def test_batched_videos_equal_one_at_a_time(dev):
    """score_videos_device (several videos per backbone pass, mixed lengths -> mixed teacher window sizes, one
    malformed video) returns bit-identical losses to scoring each video alone, in both loss modes."""
    m = build(dev, 1, 2)
    head = DINOHead(768, 4096)
    head.load_state_dict(synth_head_state_dict(2, out_dim=4096), strict=True)
    head = head.to(dev).eval()
    vids = [synth_video(s, F, smooth=0.5).to(dev) for s, F in ((1, 5), (2, 9), (3, 12), (4, 3))]
    vids.insert(2, torch.zeros(5, 3, 112, 112, device=dev))  # wrong crop -> dummy
    vids.insert(4, synth_video(9, 2, smooth=0.5).to(dev))    # shorter than the local window -> broadcast views
    dl = DINOLoss(768, 0.02, 0.3).to(dev)
    global_dummy = 6   # a dummy video yields `global_clip_size` losses
    for scorer in (None, HeadScorer(head, 0.04, 0.1)):
        with torch.no_grad():
            batch = score_videos_device(m, vids, dl, 3, 6, scorer=scorer)
            single = [score_video_device(m, v, dl, 3, 6, scorer=scorer) for v in vids]
        assert [len(b) for b in batch] == [5, 9, global_dummy, 12, 2, 3]
        for b, s in zip(batch, single):
            assert torch.equal(b, s)
    assert score_videos_device(m, [], dl) == []


def test_train_loss_and_center_update(dev):
    torch.manual_seed(0)
    out_dim, B, ncrops = 4096, 3, 4
    s, t = torch.randn(ncrops * B, out_dim), torch.randn(2 * B, out_dim)
    L = TrainDINOLoss(out_dim, ncrops, 0.04, 0.07, 3, 10).to(dev)
    c0 = torch.randn(1, out_dim) * 0.1
    L.center.copy_(c0)
    got = L(s.to(dev), t.to(dev), epoch=1)
    temp = float(L.teacher_temp_schedule[1])
    ref = O.dino_train_loss(s.double(), t.double(), c0.double(), temp, 0.1, ncrops)
    assert got.ndim == 0 and abs(float(got) - float(ref)) / float(ref) < 1e-5
    assert torch.allclose(L.center.cpu(), O.center_update(c0, [t]), atol=1e-6)


# ---- full-size properties (BASELINE config 2 / 3 shapes) -------------------------------------------------------------
def test_full_size_batch_composition_independence(dev):
    """Config 2 shape (64 clips x 16 frames x 224^2): a clip's features do not depend on what else is in the batch -
    rows of every GEMM / attention tile are independent, so results are bit-identical."""
    m = build(dev, 0, 12)
    x = synth_clips(77, 4, 16).to(dev)
    big = x.repeat(16, 1, 1, 1, 1)  # 64 clips
    with torch.no_grad():
        f_big = m(big)
        f_small = m(x)
        f_one = m(x[2:3])
    assert tuple(f_big.shape) == (64, 768) and torch.isfinite(f_big).all()
    assert torch.equal(f_big[:4], f_small) and torch.equal(f_big[60:], f_small) and torch.equal(f_small[2:3], f_one)


def test_full_size_video_invariants(dev):
    """Config 3 shape: a 4_3_30 video with F = 68 frames.  loss[0] == loss[1] and loss[-1] == loss[-2] bit-exactly
    (identical windows); windowed scoring equals scoring explicitly materialised clips."""
    m = build(dev, 0, 12)
    frames = synth_video(5, 68).to(dev)
    dl = DINOLoss(768, 0.02, 0.3).to(dev)
    with torch.no_grad():
        loss = score_video_device(m, frames, dl).cpu()
        # explicit clips for frames 0 (student [0,3), teacher [0,30)) and 40 (student [39,42), teacher [25,55))
        for i, (j, l) in ((0, (0, 0)), (40, (39, 25))):
            s = m(frames[j:j + 3].permute(1, 0, 2, 3)[None].contiguous())
            t = m(frames[l:l + 30].permute(1, 0, 2, 3)[None].contiguous())
            assert torch.equal(dl.pair_losses(s, t).cpu()[0], loss[i])
    assert loss.shape == (68,) and torch.isfinite(loss).all()
    assert loss[0] == loss[1] and loss[-1] == loss[-2]


def test_full_size_long_clip_and_knn_view_shapes(dev):
    """Config 5 (32 frames, T*N = 6 272 tokens per clip, divided and joint attention) and config 4 (8-frame 224^2 global +
    4-frame 96^2 local views grouped by MultiCropWrapper) at full model depth: finite, batch-independent features, and
    the joint kernel (K/V streaming) agrees with itself across batch compositions bit-exactly."""
    from dvst_b200.utils import MultiCropWrapper
    m = build(dev, 0, 12, T0=8)
    x = synth_clips(90, 3, 32).to(dev)
    with torch.no_grad():
        f = m(x)
        f1 = m(x[1:2])
    assert tuple(f.shape) == (3, 768) and torch.isfinite(f).all() and torch.equal(f[1:2], f1)
    mj = build(dev, 0, 2, att="joint_space_time", T0=8)
    with torch.no_grad():
        fj = mj(x[:2])
        fj1 = mj(x[1:2])
    assert torch.isfinite(fj).all() and torch.equal(fj[1:2], fj1)
    # config 4: two resolutions / frame counts through the wrapper = the same features as separate forwards
    g = synth_clips(91, 4, 8).to(dev)
    l = synth_clips(92, 4, 4, 96, 96).to(dev)
    w = MultiCropWrapper(m, torch.nn.Identity())
    with torch.no_grad():
        out = w([g, l])
        fg, fl_ = m(g), m(l)
    assert tuple(out.shape) == (8, 768) and torch.equal(out[:4], fg) and torch.equal(out[4:], fl_)


def test_pipeline_tool_single_gpu(dev, tmp_path):
    """tools/pipeline_demo.py: uint8 upload -> fused preprocess -> batched scoring -> gather -> JSON -> device selection,
    with its own --check (every video recomputed alone: bit-identical losses, identical index lists).  The same command
    under torchrun --nproc-per-node 2 (NCCL) writes the identical file (verified on 2 B200s; not repeated here)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "pipeline_demo.py"), "--videos", "6", "--depth", "2",
                          "--out", str(tmp_path / "loss.json"), "--check"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "pipeline ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    d = json.load(open(tmp_path / "loss.json"))
    assert len(d) == 6 and len(d["video_0000"]) == 2    # the 2-frame video: the reference's broadcast views, 2 losses


def test_forward_is_cuda_graph_capturable(dev):
    """The C ABI promises no device synchronisation and no device allocation inside a call (include/dvst.h): a whole
    forward (146 launches for ViT-B) can therefore be captured into a CUDA graph and replayed - the way to run small,
    launch-bound batches.  The replay on new input equals the eager call bit for bit."""
    m = build(dev, 1, 2)
    x = synth_clips(70, 2, 3).to(dev)
    x2 = synth_clips(71, 2, 3).to(dev)
    with torch.no_grad():
        ref2 = m(x2).clone()
        m(x)                                            # weights packed, workspace allocated, kernel attributes set
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m(x)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m(x)
        x.copy_(x2)
        g.replay()
        torch.cuda.synchronize()
    assert torch.equal(out, ref2)


def test_nccl_world2_extract_features_update_center_gather_scores(dev):
    """The path's three collectives under NCCL with the real model on two GPUs (tools/nccl_check.py): extract_features'
    all-gathers (eval_knn.py:111-132), the training loss' center all-reduce (train_ssl.py:673-691) and the gather of
    per-video score vectors, each compared with the single-GPU result.  Skipped on a one-GPU box (the driver's GPU tier);
    run with `gpurun --gpus 2`."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541",
                          os.path.join(root, "tools", "nccl_check.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "nccl ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_training_forward_has_a_pytorch_backward(dev):
    """student forward through MultiCropWrapper -> DINOHead -> training DINOLoss with
    loss.backward() (train_ssl.py:483-540).  Forward values are the kernels'; gradients come from the PyTorch
    recomputation in dvst_b200/autograd.py and must equal torch.autograd through the oracle (fp32) on the same inputs."""
    from dvst_b200.utils import MultiCropWrapper
    torch.manual_seed(0)
    sd = synth_backbone_state_dict(6, depth=2, num_frames=8)
    hsd = synth_head_state_dict(3, out_dim=4096)
    m = VisionTransformer(img_size=224, num_classes=0, depth=2, qkv_bias=True,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
    m.load_state_dict(sd, strict=True)
    head = DINOHead(768, 4096)
    head.load_state_dict(hsd, strict=True)
    student = MultiCropWrapper(m, head).to(dev).train()
    B = 2
    views = [synth_clips(300 + i, B, 2, 224, 224).to(dev) for i in range(2)] + \
            [synth_clips(310 + i, B, 2, 96, 96).to(dev) for i in range(2)]       # 2 global (224^2) + 2 local (96^2) views
    teacher_out = torch.randn(2 * B, 4096, generator=torch.Generator().manual_seed(9)).to(dev)
    L = TrainDINOLoss(4096, 4, 0.04, 0.07, 3, 10).to(dev)
    c0 = L.center.clone()
    out = student(views)
    assert out.requires_grad and tuple(out.shape) == (4 * B, 4096)
    loss = L(out, teacher_out, epoch=1)
    loss.backward()
    got = {n: p.grad.detach().cpu() for n, p in student.named_parameters() if p.grad is not None}
    assert "backbone.blocks.0.temporal_fc.weight" in got and "head.mlp.0.weight" in got and "head.last_layer.weight_v" in got

    # oracle: the same computation in fp32 torch on the CPU, differentiated by torch.autograd
    osd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ohd = {k: v.clone().requires_grad_(k != "last_layer.weight_g") for k, v in hsd.items()}
    feats = torch.cat([O.forward_features(osd, torch.cat([v.cpu() for v in views[:2]])),
                       O.forward_features(osd, torch.cat([v.cpu() for v in views[2:]]))])
    o_out = O.dino_head(ohd, feats)
    o_loss = O.dino_train_loss(o_out, teacher_out.cpu(), c0.cpu(), float(L.teacher_temp_schedule[1]), 0.1, 4)
    o_loss.backward()
    assert abs(float(loss) - float(o_loss)) / float(o_loss) < 2e-3          # forward: kernels (bf16 operands) vs fp32
    worst = 0.0
    for k, v in osd.items():
        g = got["backbone." + k]
        worst = max(worst, ((g - v.grad).norm() / (v.grad.norm() + 1e-12)).item())
    for k, v in ohd.items():
        if v.requires_grad:
            g = got["head." + k]
            worst = max(worst, ((g - v.grad).norm() / (v.grad.norm() + 1e-12)).item())
    # the backward runs on the kernels' student output (bf16-level forward difference feeds softmax(s / tau)): a few 1e-3
    assert worst < 2e-2, worst


@pytest.mark.parametrize("prec", ["bf16", "high"])
def test_windows_from_decoded_uint8_frames_equal_the_two_step_path(dev, prec):
    """dvst_forward_windows_u8 (sub-sampling, / 255, mean / std, centre crop, window gather and patch extraction in one
    pass over the decoded uint8 frames - datasets_custom/dino_loss_loader.py:80-92,126-171 + models/timesformer.py:201-208)
    gives the same BITS as preprocess_frames() followed by the window forward, and agrees with the CPU oracle of the loader
    front half + backbone."""
    from dvst_b200.dino_similarity import _forward_windows, forward_windows_u8
    m = build(dev, 3, 2, "divided_space_time", prec)
    g = torch.Generator().manual_seed(77)
    for (n_frames, hf, wf, rate, crop, T, starts) in [(22, 240, 320, 2, 224, 4, [0, 3, 7, 2]), (9, 224, 224, 1, 224, 3, [0, 6]),
                                                      (16, 130, 171, 4, 96, 2, [0, 1, 2]), (8, 256, 256, 1, 224, 8, [0])]:
        raw = torch.randint(0, 256, (n_frames, hf, wf, 3), generator=g, dtype=torch.uint8)
        with torch.no_grad():
            frames = preprocess_frames(raw.to(dev), rate, crop_size=crop)
            two = _forward_windows(m, frames, starts, T, 3)
            one = forward_windows_u8(m, raw.to(dev), starts, T, 3, sampling_rate=rate, crop_size=crop)
        assert torch.equal(one, two)
    # against the CPU statement of the loader front half + the fp32 oracle backbone (last case: one 8-frame clip)
    sd = synth_backbone_state_dict(3, depth=2, num_frames=8)
    ref = O.forward_features(sd, O.preprocess_frames(raw, 1, crop_size=224).permute(1, 0, 2, 3)[None])
    assert _rel(one, ref) < (FEAT_TOL if prec == "bf16" else 1e-4)
    with pytest.raises(_lib.DvstError):
        forward_windows_u8(m, raw, [0], 8)                       # CPU input: no fallback
    with pytest.raises(_lib.DvstError):
        forward_windows_u8(m, raw.to(dev), [0], 9)               # window longer than the video
