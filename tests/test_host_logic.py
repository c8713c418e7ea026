"""Host-side mirror of the reference interfaces (no GPU): state_dict key parity, window / selection logic vs the
reference goldens, JSON export format, sharding, and 'fails loudly on CPU'."""
import json
import os
import types
from functools import partial

import pytest
import torch

import dvst_b200  # noqa: F401
from dvst_b200 import _lib
from dvst_b200.dino_similarity import DINOLoss, export_loss, export_losses, preprocess_frames, window_starts
from dvst_b200.frame_selection import select_frames_adaptive, select_frames_uniform
from dvst_b200.models import TimeSformer, VisionTransformer, get_vit_base_patch16_224, vit_base_patch16_224
from dvst_b200.parallel import shard_videos, video_cost
from dvst_b200.testing import synth_backbone_state_dict, synth_head_state_dict
from dvst_b200.utils import MultiCropWrapper
from dvst_b200.vision_transformer import DINOHead


def _cfg(att="divided_space_time"):
    return types.SimpleNamespace(DATA=types.SimpleNamespace(TRAIN_CROP_SIZE=224, NUM_FRAMES=8),
                                 MODEL=types.SimpleNamespace(NUM_CLASSES=400),
                                 TIMESFORMER=types.SimpleNamespace(ATTENTION_TYPE=att, PRETRAINED_MODEL=""))


def test_factory_matches_reference_interface():
    m = get_vit_base_patch16_224(_cfg(), no_head=False)
    assert sum(p.numel() for p in m.parameters()) == 121566352  # 121.57 M
    assert m.embed_dim == m.num_features == 768 and m.num_patches == 196
    assert m.attention_type == "divided_space_time" and len(m.blocks) == 12
    # reference init quirk: temporal_fc of ALL blocks and time_embed are zero
    assert all(float(b.temporal_fc.weight.detach().abs().sum()) == 0 for b in m.blocks)
    assert float(m.time_embed.detach().abs().sum()) == 0
    res = m.load_state_dict(synth_backbone_state_dict(0, num_classes=400), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert get_vit_base_patch16_224(_cfg(), no_head=True).head is None
    # checkpoint loading path of the drivers: strict=False after the "backbone." strip
    ck = {"backbone." + k: v for k, v in synth_backbone_state_dict(0).items()}
    ren = {k[len("backbone."):]: v for k, v in ck.items() if k.startswith("backbone.")}
    assert set(m.load_state_dict(ren, strict=False).missing_keys) == {"head.weight", "head.bias"}


@pytest.mark.parametrize("att", ["joint_space_time", "space_only"])
def test_other_attention_types_keys(att):
    m = VisionTransformer(num_classes=0, depth=2, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                          attention_type=att)
    m.load_state_dict(synth_backbone_state_dict(0, depth=2, attention_type=att), strict=True)


def test_wrappers_and_unsupported_configs():
    assert isinstance(TimeSformer(num_classes=0).model, VisionTransformer)
    assert isinstance(vit_base_patch16_224(_cfg()).model, VisionTransformer)
    with pytest.raises(ValueError):
        VisionTransformer(embed_dim=384, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    with pytest.raises(ValueError):
        VisionTransformer(qkv_bias=True)  # default LayerNorm eps 1e-5 is not what the reference builds


def test_cpu_input_fails_loudly():
    m = VisionTransformer(num_classes=0, depth=1, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6))
    with pytest.raises(_lib.DvstError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 2, 224, 224))
    with pytest.raises(_lib.DvstError, match="no CPU fallback"):
        DINOLoss(768, 0.02, 0.3)(torch.zeros(768), torch.zeros(768))
    with pytest.raises(_lib.DvstError, match="no CPU fallback"):
        DINOHead(768, 1024)(torch.zeros(2, 768))


def test_dino_head_keys_and_loss_buffer():
    h = DINOHead(768, 65536)
    assert sum(p.numel() for p in h.parameters()) == 23138560  # 23.14 M (incl. weight_g)
    res = h.load_state_dict(synth_head_state_dict(2), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert float(h.last_layer.weight_g.min()) == 1.0 and not h.last_layer.weight_g.requires_grad
    L = DINOLoss(out_dim=768, teacher_temp=0.02, student_temp=0.3)
    assert tuple(L.center.shape) == (1, 768) and "center" in L.state_dict()


def test_window_rule_and_selection_vs_reference_goldens(golden_dir):
    W = json.load(open(os.path.join(golden_dir, "windows.json")))
    for F, c in W.items():
        G, ls, gs = window_starts(int(F))
        assert G == c["G"] and [list(p) for p in zip(ls, gs)] == [list(p) for p in c["starts"]]
        # only F-2 student and max(F-G+1, 1) teacher windows are distinct (SURVEY finding 5)
        assert len(set(ls)) == int(F) - 2 and len(set(gs)) == max(int(F) - G + 1, 1)
    S = json.load(open(os.path.join(golden_dir, "selection.json")))
    for name, c in S.items():
        for N in (8, 16):
            assert select_frames_adaptive(c["loss"], c["n_unsampled"], 4, N) == c[f"idx{N}"], name
    assert select_frames_uniform(75, 8) == [0, 9, 18, 27, 36, 45, 54, 63]


def test_export_loss_format(tmp_path):
    p = str(tmp_path / "loss.json")
    export_loss([1.0, 2.5], "/data/vids/abc_1_2.avi", p)
    export_loss([3.0], "/x/def.mp4", p)
    export_loss([9.0], "abc_1_2.avi", p)  # read-modify-write: later result replaces the earlier one
    assert open(p).read() == '{"abc_1_2": [9.0], "def": [3.0]}'


def test_export_losses_equals_sequential_export_loss(tmp_path):
    """One read-modify-write for a batch of videos leaves the same file as the reference's per-video export_loss."""
    a, b = str(tmp_path / "a.json"), str(tmp_path / "b.json")
    items = [("/d/v1.avi", [1.0, 2.0]), ("/d/v2.mp4", [0.5]), ("/e/v1.avi", [7.0]), ("v3.avi", [3.25, 4.5, 1e-3])]
    export_loss([8.0], "old.avi", a)
    export_loss([8.0], "old.avi", b)
    for path, loss in items[:2]:
        export_loss(loss, path, a)
    for path, loss in items[2:]:
        export_loss(loss, path, a)
    export_losses(dict(items[:2]), b)
    export_losses(dict(items[2:]), b)
    assert open(a).read() == open(b).read()


def test_preprocess_frames():
    import dvst_oracle as O
    raw = torch.randint(0, 256, (10, 240, 320, 3), dtype=torch.uint8)
    f = O.preprocess_frames(raw, 4)
    assert tuple(f.shape) == (3, 3, 224, 224)
    ref = ((raw[4, 8 + 5, 48 + 7, 1].float() / 255.0) - 0.45) / 0.225
    assert abs(float(f[1, 1, 5, 7]) - float(ref)) < 1e-6
    with pytest.raises(_lib.DvstError):       # the product has no CPU / PyTorch path for it
        preprocess_frames(raw, 4)


def test_sharding_is_balanced_and_complete():
    lens = [30, 68, 75, 11, 450, 38, 38, 53, 2, 300, 120, 64]
    for world in (1, 2, 4, 8):
        parts = shard_videos(lens, world)
        assert sorted(i for p in parts for i in p) == list(range(len(lens)))
        loads = [sum(video_cost(lens[i]) for i in p) for p in parts]
        assert max(loads) <= sum(loads) / world + max(video_cost(f) for f in lens)


def test_multicrop_wrapper_groups_by_resolution():
    calls = []

    class Fake(torch.nn.Module):
        def forward(self, x):
            calls.append(tuple(x.shape))
            return x.flatten(1)[:, :4]
    w = MultiCropWrapper(Fake(), torch.nn.Identity())
    views = [torch.zeros(2, 3, 8, 224, 224)] * 2 + [torch.zeros(2, 3, 8, 96, 96)] * 3
    out = w(views)
    assert calls == [(4, 3, 8, 224, 224), (6, 3, 8, 96, 96)] and tuple(out.shape) == (10, 4)


def test_product_code_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package, include/ or the root shim may import or execute it
    (bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() are the two allowed callers)."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "dino-video-summarization-transformer_b200", "**", "*.py"), recursive=True)
    files += glob.glob(os.path.join(root, "dino-video-summarization-transformer_b200", "csrc", "*.*"))
    files += [os.path.join(root, "dvst_b200.py")] + glob.glob(os.path.join(root, "include", "*.h"))
    pat = re.compile(r"dvst_oracle|ref_shim|make_golden|[\"'/]oracle[\"'/]|import oracle|from oracle")
    bad = [f for f in files if not f.endswith(".so") and pat.search(open(f, errors="ignore").read())]
    assert len(files) > 15 and not bad, bad


def test_load_config_and_decoded_dataset(tmp_path, monkeypatch):
    """config.load_config mirrors utils/parser.py:65-90 for the fields the scoring path reads (defaults, YAML, --opts);
    config.DecodedVideoDataset mirrors the list / decode half of DinoLossLoader (dino_loss_loader.py:40-82)."""
    import types
    from dvst_b200.config import DecodedVideoDataset, load_config
    yml = tmp_path / "cfg.yaml"
    yml.write_text("DATA:\n  NUM_FRAMES: 8\n  TRAIN_CROP_SIZE: 224\nTIMESFORMER:\n  ATTENTION_TYPE: 'divided_space_time'\n"
                   "TEST:\n  NUM_ENSEMBLE_VIEWS: 1\nMODEL:\n  NUM_CLASSES: 400\n")
    (tmp_path / "test.csv").write_text("a.avi 0\nsub/b.avi 3\n")
    args = types.SimpleNamespace(cfg_file=str(yml), opts=["DATA.PATH_TO_DATA_DIR", str(tmp_path), "DATA.PATH_PREFIX",
                                                           "/videos", "DATA.USE_FLOW", "False"],
                                 pretrained_weights="x.pth", num_workers=4, output_dir="out")
    cfg = load_config(args)
    assert cfg.DATA.PATH_TO_DATA_DIR == str(tmp_path) and cfg.DATA.PATH_PREFIX == "/videos" and cfg.DATA.USE_FLOW is False
    assert cfg.DATA.MEAN == [0.45, 0.45, 0.45] and cfg.DATA.STD == [0.225, 0.225, 0.225] and cfg.DATA.NUM_FRAMES == 8
    assert cfg.TEST.NUM_ENSEMBLE_VIEWS == 1 and cfg.MODEL.NUM_CLASSES == 400 and cfg.OUTPUT_DIR == "out"
    assert cfg.TIMESFORMER.ATTENTION_TYPE == "divided_space_time" and cfg.TIMESFORMER.PRETRAINED_MODEL == ""
    assert load_config(cfg) is cfg                       # an already resolved config passes through
    assert load_config(types.SimpleNamespace(cfg_file=None, opts=None)).TEST.NUM_ENSEMBLE_VIEWS == 10   # defaults.py:96
    ds = DecodedVideoDataset(cfg, "test")
    assert len(ds) == 2 and ds._path_to_videos == ["/videos/a.avi", "/videos/sub/b.avi"] and ds._labels == [0, 3]
    import torchvision.io as io
    monkeypatch.setattr(io, "read_video", lambda path, pts_unit="sec": (torch.zeros(5, 4, 4, 3, dtype=torch.uint8), None, None),
                        raising=False)
    video, path = ds[1]
    assert video.shape == (5, 4, 4, 3) and path == "/videos/sub/b.avi"
    import pytest
    with pytest.raises(FileNotFoundError):
        DecodedVideoDataset(cfg, "val")


def test_short_video_clips_match_reference_loader_golden(golden_dir):
    import json
    from dvst_b200.dino_similarity import short_video_clips
    import dvst_oracle as O
    W = json.load(open(os.path.join(golden_dir, "windows_short.json")))
    for F in (1, 2):
        si, ti = short_video_clips(F, 3)
        assert O.short_video_clips(F, 3) == (si, ti) and len(W[str(F)]) == F
        assert all(p["student"] == si and p["teacher"] == ti for p in W[str(F)])


def test_fused_temporal_kernel_tile_geometry():
    """Host logic of csrc/gemm_tattn_sm100.cu (no device work): for every T <= 64 a CTA tile holds whole sequences, and the
    key block of every 16-row query tile (first row of the first sequence it touches, moved down when it would pass row 128)
    contains every sequence those 16 rows belong to - the property the in-kernel sequence mask relies on."""
    import ctypes
    from dvst_b200 import _lib
    L = _lib.lib()
    for T in range(1, 65):
        R, KB = ctypes.c_int(), ctypes.c_int()
        assert L.dvst_temporal_tile_geometry(T, ctypes.byref(R), ctypes.byref(KB)) == 0
        R, KB = R.value, KB.value
        assert R % T == 0 and 0 < R <= 128 and KB in (16, 32, 64)
        R0 = (128 // T) * T      # as many whole sequences as fit, unless a query tile would then span more than 64 rows
        span0 = max(((min(r0 + 15, R0 - 1) // T) + 1) * T - (r0 // T) * T for r0 in range(0, R0, 16))
        assert R == (R0 if span0 <= 64 else T)
        for r0 in range(0, R, 16):
            key0 = min((r0 // T) * T, 128 - KB)
            for r in range(r0, min(r0 + 16, R)):
                s0 = (r // T) * T
                assert key0 <= s0 and s0 + T <= key0 + KB, (T, r0, r)
        if T & (T - 1) == 0:
            assert R == 128 and KB == max(16, T)       # power-of-two T: no masking needed from T = 16 up
    R, KB = ctypes.c_int(), ctypes.c_int()
    for T in (0, 65, 1000):
        assert L.dvst_temporal_tile_geometry(T, ctypes.byref(R), ctypes.byref(KB)) != 0
        assert b"not covered" in L.dvst_last_error()


@pytest.mark.parametrize("prec", ["bf16", "high"])
def test_weight_packing_folds_layernorm_into_the_consumers(prec):
    """Packing (no GPU needed: the handle only records pointers) registers, for the three Linears that consume a LayerNorm
    (temporal qkv, qkv, fc1), W' = gamma * W in the GEMM operand format of the precision (bf16, or the [hi | lo] pair),
    colsum(W') and b + W beta - the tensors of the deferred LayerNorm (csrc/gemm_sm100.cu) - and folds
    temporal_fc(temporal_attn.proj(.)) into one 768 x 768 Linear (models/timesformer.py:83,133)."""
    m = VisionTransformer(num_classes=0, depth=1, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                          num_frames=8, precision=prec)
    m.load_state_dict(synth_backbone_state_dict(3, depth=1, num_frames=8), strict=True)
    m._ensure_packed(torch.device("cpu"))
    pk, blk = m._packed, m.blocks[0]
    k = 2 if prec == "high" else 1
    for name, lin, norm in (("temporal_attn.qkv", blk.temporal_attn.qkv, blk.temporal_norm1),
                            ("attn.qkv", blk.attn.qkv, blk.norm1), ("mlp.fc1", blk.mlp.fc1, blk.norm2)):
        wf, sf, bf = (pk[f"blocks.0.{name}.{s}"] for s in ("wf", "sf", "bf"))
        W, g, beta = lin.weight.detach().double(), norm.weight.detach().double(), norm.bias.detach().double()
        assert wf.dtype == torch.bfloat16 and wf.shape == (W.shape[0], 768 * k)
        w_eff = wf[:, :768].double() + (wf[:, 768:].double() if k == 2 else 0.0)      # what the tensor core multiplies by
        rel = ((w_eff - W * g[None]).norm() / (W * g[None]).norm()).item()
        assert rel < (2e-5 if k == 2 else 4e-3), rel                                    # 16 vs 8 significand bits
        assert torch.allclose(sf.double(), w_eff.sum(1), rtol=0, atol=1e-5)             # colsum of the ROUNDED operand
        assert torch.allclose(bf.double(), lin.bias.detach().double() + W @ beta, rtol=0, atol=1e-6)
    # the head-permuted copies of the fused temporal kernel exist in bf16 precision only
    assert ("blocks.0.temporal_attn.qkv.wfp" in pk) == (prec == "bf16")
    wfc, wp = blk.temporal_fc.weight.detach().double(), blk.temporal_attn.proj.weight.detach().double()
    w = pk["blocks.0.temporal_projfc.w"]
    w_eff = w[:, :768].double() + (w[:, 768:].double() if k == 2 else 0.0)
    assert ((w_eff - wfc @ wp).norm() / (wfc @ wp).norm()).item() < (2e-5 if k == 2 else 4e-3)
