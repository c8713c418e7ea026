"""Seeded synthetic weights / inputs shared by tests and bench.py (no datasets or checkpoints exist offline).

Plain random-init of the reference leaves the temporal path dead (`temporal_fc` and `time_embed` are zero after
models/timesformer.py:234,255-263) and all biases zero, which would let a broken kernel pass parity.  These
generators therefore randomise EVERY tensor (biases, LayerNorm affine, temporal_fc, time_embed) and use a larger
std on the qkv projections so the attention softmaxes are far from uniform.  All draws come from a CPU
torch.Generator, so the same seed gives bit-identical tensors on every machine.
"""
import torch

EMBED = 768


def _n(g, *shape, std=0.02, mean=0.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std + mean


def synth_backbone_state_dict(seed=0, depth=12, num_frames=8, img_size=224, attention_type="divided_space_time",
                              num_classes=0, hard=False):
    """State dict with the reference VisionTransformer's keys (models/timesformer.py:212-252).
    hard=True mimics what trained ViT checkpoints do to the numerics: sharp attention (large q / k projections), LayerNorm
    gains up to ~3, a residual-stream channel carrying a massive activation and a common offset on every token (rows whose
    mean is several times their spread) - the regimes that stress bf16 operands, the deferred LayerNorm and the one-pass
    softmax."""
    sd = _synth_backbone(seed, depth, num_frames, img_size, attention_type, num_classes)
    if hard:
        g = torch.Generator().manual_seed(seed + 7919)
        for k in list(sd):
            if k.endswith("qkv.weight"):
                sd[k] = sd[k] * 1.5                                   # logits ~2.3x larger: sharp attention rows
            elif ".norm" in k and k.endswith(".weight") or k.endswith("temporal_norm1.weight"):
                sd[k] = sd[k] * (1.0 + 2.0 * torch.rand(sd[k].shape, generator=g))
        sd["patch_embed.proj.bias"] = sd["patch_embed.proj.bias"] + 2.0   # common offset on every token
        sd["patch_embed.proj.bias"][5] += 40.0                            # a massive-activation channel
        for i in range(depth):
            sd[f"blocks.{i}.mlp.fc2.bias"] = sd[f"blocks.{i}.mlp.fc2.bias"] + 0.3
    return sd


def _synth_backbone(seed, depth, num_frames, img_size, attention_type, num_classes):
    g = torch.Generator().manual_seed(seed)
    P = (img_size // 16) ** 2
    sd = {
        "cls_token": _n(g, 1, 1, EMBED),
        "pos_embed": _n(g, 1, P + 1, EMBED),
        "patch_embed.proj.weight": _n(g, EMBED, 3, 16, 16),
        "patch_embed.proj.bias": _n(g, EMBED),
        "norm.weight": _n(g, EMBED, std=0.05, mean=1.0),
        "norm.bias": _n(g, EMBED),
    }
    if attention_type != "space_only":
        sd["time_embed"] = _n(g, 1, num_frames, EMBED)
    for i in range(depth):
        p = f"blocks.{i}."
        norms = ["norm1", "norm2"] + (["temporal_norm1"] if attention_type == "divided_space_time" else [])
        for n in norms:
            sd[p + n + ".weight"] = _n(g, EMBED, std=0.05, mean=1.0)
            sd[p + n + ".bias"] = _n(g, EMBED)
        attns = ["attn"] + (["temporal_attn"] if attention_type == "divided_space_time" else [])
        for a in attns:
            sd[p + a + ".qkv.weight"] = _n(g, 3 * EMBED, EMBED, std=0.04)
            sd[p + a + ".qkv.bias"] = _n(g, 3 * EMBED)
            sd[p + a + ".proj.weight"] = _n(g, EMBED, EMBED)
            sd[p + a + ".proj.bias"] = _n(g, EMBED)
        if attention_type == "divided_space_time":
            sd[p + "temporal_fc.weight"] = _n(g, EMBED, EMBED)
            sd[p + "temporal_fc.bias"] = _n(g, EMBED)
        sd[p + "mlp.fc1.weight"] = _n(g, 4 * EMBED, EMBED)
        sd[p + "mlp.fc1.bias"] = _n(g, 4 * EMBED)
        sd[p + "mlp.fc2.weight"] = _n(g, EMBED, 4 * EMBED)
        sd[p + "mlp.fc2.bias"] = _n(g, EMBED)
    if num_classes > 0:
        sd["head.weight"] = _n(g, num_classes, EMBED)
        sd["head.bias"] = _n(g, num_classes)
    return sd


def synth_head_state_dict(seed=2, in_dim=EMBED, out_dim=65536, hidden_dim=2048, bottleneck_dim=256):
    """State dict with DINOHead's keys (vision_transformer.py:248-271): mlp.{0,2,4}, last_layer.weight_{g,v}."""
    g = torch.Generator().manual_seed(seed)
    return {
        "mlp.0.weight": _n(g, hidden_dim, in_dim), "mlp.0.bias": _n(g, hidden_dim),
        "mlp.2.weight": _n(g, hidden_dim, hidden_dim), "mlp.2.bias": _n(g, hidden_dim),
        "mlp.4.weight": _n(g, bottleneck_dim, hidden_dim), "mlp.4.bias": _n(g, bottleneck_dim),
        "last_layer.weight_g": torch.ones(out_dim, 1),
        "last_layer.weight_v": _n(g, out_dim, bottleneck_dim),
    }


def synth_clips(seed, B, T, H=224, W=224):
    """Clips shaped like the reference's model input [B,3,T,H,W] (mean/std-normalised pixels ~ N(0,1))."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, T, H, W, generator=g, dtype=torch.float32)


def synth_video(seed, F, H=224, W=224, smooth=0.7):
    """A [F,3,H,W] 'video': temporally correlated frames so neighbouring windows differ smoothly."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randn(F, 3, H, W, generator=g, dtype=torch.float32)
    for i in range(1, F):
        frames[i] = smooth * frames[i - 1] + (1.0 - smooth ** 2) ** 0.5 * frames[i]
    return frames


def synth_knn_features(seed, n_train, n_test, dim=768, num_classes=10, noise=1.0):
    """L2-normalised class-clustered features + int64 labels shaped like eval_knn's inputs (eval_knn.py:69-74):
    (train_features [n_train, dim], train_labels, test_features [n_test, dim], test_labels)."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(num_classes, dim, generator=g)
    out = []
    for n in (n_train, n_test):
        labels = torch.randint(0, num_classes, (n,), generator=g, dtype=torch.int64)
        f = centres[labels] + noise * torch.randn(n, dim, generator=g)
        out += [torch.nn.functional.normalize(f, dim=1, p=2), labels]
    return tuple(out)
