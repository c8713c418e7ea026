"""Drop-in for the reference's models/timesformer.py (divided / joint / space-only TimeSformer ViT-B/16).

Same class names, constructor signatures, attributes and state_dict keys as the reference
(models/timesformer.py:212-263, 379-426, 592-609), so SVT checkpoints load with
`load_state_dict(strict=False)` after the usual "backbone." strip (dino_similarity.py:32-37).  The modules below are
parameter containers only: `forward` packs the weights once (bf16, or bf16 hi/lo pairs in `precision="high"`) and
calls `dvst_forward_features` in csrc/libdvst.so, which runs the whole forward as hand-written sm_100a kernels on a
resident [B,N,T,D] token layout.  There is no PyTorch / CPU fallback: a CPU input or a missing library raises.

The reference's scoring / feature-extraction callers run under torch.no_grad(); when gradients are recorded (training a
student) the forward still runs on the kernels and the backward is PyTorch (dvst_b200/autograd.py).
"""
import ctypes
import math
import os
from functools import partial

import torch
import torch.nn as nn

from .. import _lib

_ATTN = {"divided_space_time": 0, "joint_space_time": 1, "space_only": 2}
_PRECISION = {"bf16": 0, "high": 1}


class _Config(ctypes.Structure):
    _fields_ = [("depth", ctypes.c_int), ("num_frames", ctypes.c_int), ("pos_grid", ctypes.c_int),
                ("attention_type", ctypes.c_int), ("precision", ctypes.c_int)]


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """Same distribution as models/vit_utils.py:trunc_normal_ (truncation at +-2 in absolute units)."""
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class Mlp(nn.Module):
    """Parameter container with the reference's names (models/timesformer.py:36-44)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)


class Attention(nn.Module):
    """Parameter container (models/timesformer.py:55-66); the math runs in csrc/attention.cu."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., with_qkv=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.with_qkv = with_qkv
        if with_qkv:
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.proj = nn.Linear(dim, dim)


class Block(nn.Module):
    """Parameter container (models/timesformer.py:90-114)."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0.1, act_layer=nn.GELU, norm_layer=nn.LayerNorm, attention_type='divided_space_time'):
        super().__init__()
        assert attention_type in ['divided_space_time', 'space_only', 'joint_space_time']
        self.attention_type = attention_type
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale)
        if attention_type == 'divided_space_time':
            self.temporal_norm1 = norm_layer(dim)
            self.temporal_attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale)
            self.temporal_fc = nn.Linear(dim, dim)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)


class PatchEmbed(nn.Module):
    """Parameter container (models/timesformer.py:188-201)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        img_size, patch_size = to_2tuple(img_size), to_2tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class VisionTransformer(nn.Module):
    """Vision Transformer with the reference's constructor (models/timesformer.py:215-218) plus `precision`:
    "bf16" (bf16 operands, fp32 accumulate / residual / LayerNorm / softmax) or "high" (2-term bf16 split in every
    GEMM, ~16 mantissa bits, 3x the tensor work; what reference-exact mode-R scoring needs, SURVEY.md finding 3)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0.1, hybrid_backbone=None, norm_layer=nn.LayerNorm, num_frames=8,
                 attention_type='divided_space_time', dropout=0., precision=None):
        super().__init__()
        if (patch_size, in_chans, embed_dim, num_heads, int(mlp_ratio)) != (16, 3, 768, 12, 4) or not qkv_bias:
            raise ValueError("the sm_100a kernels are specialised for ViT-B/16 (patch 16, 3 channels, width 768, "
                             "12 heads, mlp_ratio 4, qkv_bias=True) - the only configuration the reference builds "
                             "(models/timesformer.py:384-388,409-413,594-598)")
        if qk_scale is not None and abs(qk_scale - 0.125) > 1e-12:
            raise ValueError("qk_scale must be None (= 64 ** -0.5)")
        if attention_type not in _ATTN:
            raise ValueError(attention_type)
        eps = getattr(norm_layer(8), "eps", None)
        if eps is None or abs(eps - 1e-6) > 1e-12:
            raise ValueError("norm_layer must be partial(nn.LayerNorm, eps=1e-6) as in the reference factory")
        precision = precision or os.environ.get("DVST_PRECISION", "bf16")
        if precision not in _PRECISION:
            raise ValueError(f"precision must be one of {list(_PRECISION)}")
        self.precision = precision
        self.attention_type = attention_type
        self.depth = depth
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        if attention_type != 'space_only':
            self.time_embed = nn.Parameter(torch.zeros(1, num_frames, embed_dim))
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  norm_layer=norm_layer, attention_type=attention_type) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()

        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)
        if attention_type == 'divided_space_time':
            # The reference's loop (models/timesformer.py:255-263) is meant to skip block 0 but its first match is the
            # ModuleList itself, so temporal_fc of ALL blocks ends up zero; reproduced here.
            for blk in self.blocks:
                nn.init.constant_(blk.temporal_fc.weight, 0)
                nn.init.constant_(blk.temporal_fc.bias, 0)

        self._handle = None
        self._packed = {}
        self._pack_key = None
        self._workspace = None

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token', 'time_embed'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    # ------------------------------------------------------------------------------------------------------------------
    # weight packing: fp32 parameters -> device buffers in the layout dvst_set_weight documents (include/dvst.h)
    # ------------------------------------------------------------------------------------------------------------------
    def _backbone_params(self):
        return [p for n, p in self.named_parameters() if not n.startswith("head.")]

    def _current_key(self, device):
        ps = self._backbone_params()
        return (str(device), self.precision, tuple((p.data_ptr(), p._version) for p in ps))

    def _gemm_weight(self, w, device):
        w = w.detach().to(device=device, dtype=torch.float32).reshape(w.shape[0], -1)
        if self.precision == "high":
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            return torch.cat([hi, lo], dim=1).contiguous()
        return w.to(torch.bfloat16).contiguous()

    def _ensure_packed(self, device):
        key = self._current_key(device)
        if key == self._pack_key and self._handle is not None:
            return
        L = _lib.lib()
        if self._handle is not None:
            L.dvst_destroy(self._handle)
            self._handle = None
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        pk = {}
        pk["patch_embed.w"] = self._gemm_weight(self.patch_embed.proj.weight, device)
        pk["patch_embed.b"] = f32(self.patch_embed.proj.bias)
        pk["cls_pos0"] = f32(self.cls_token[0, 0] + self.pos_embed[0, 0])
        pk["pos_embed"] = f32(self.pos_embed[0])
        if self.attention_type != 'space_only':
            pk["time_embed"] = f32(self.time_embed[0])
        pk["norm.w"], pk["norm.b"] = f32(self.norm.weight), f32(self.norm.bias)
        for i, blk in enumerate(self.blocks):
            p = f"blocks.{i}."
            names = ["norm1", "norm2"] + (["temporal_norm1"] if self.attention_type == 'divided_space_time' else [])
            for n in names:
                pk[p + n + ".w"], pk[p + n + ".b"] = f32(getattr(blk, n).weight), f32(getattr(blk, n).bias)
            if self.attention_type == 'divided_space_time':
                pk[p + "temporal_attn.qkv.w"] = self._gemm_weight(blk.temporal_attn.qkv.weight, device)
                pk[p + "temporal_attn.qkv.b"] = f32(blk.temporal_attn.qkv.bias)
                # temporal_fc(temporal_attn.proj(a)) has no non-linearity in between (models/timesformer.py:83,133):
                # fold the two 768x768 linears into one (composed in fp64).
                wf = blk.temporal_fc.weight.detach().to(device).double()
                wp = blk.temporal_attn.proj.weight.detach().to(device).double()
                bp = blk.temporal_attn.proj.bias.detach().to(device).double()
                bf = blk.temporal_fc.bias.detach().to(device).double()
                pk[p + "temporal_projfc.w"] = self._gemm_weight((wf @ wp).float(), device)
                pk[p + "temporal_projfc.b"] = (wf @ bp + bf).float().contiguous()
            pk[p + "attn.qkv.w"] = self._gemm_weight(blk.attn.qkv.weight, device)
            pk[p + "attn.qkv.b"] = f32(blk.attn.qkv.bias)
            pk[p + "attn.proj.w"] = self._gemm_weight(blk.attn.proj.weight, device)
            pk[p + "attn.proj.b"] = f32(blk.attn.proj.bias)
            pk[p + "mlp.fc1.w"] = self._gemm_weight(blk.mlp.fc1.weight, device)
            pk[p + "mlp.fc1.b"] = f32(blk.mlp.fc1.bias)
            pk[p + "mlp.fc2.w"] = self._gemm_weight(blk.mlp.fc2.weight, device)
            pk[p + "mlp.fc2.b"] = f32(blk.mlp.fc2.bias)
        # Deferred LayerNorm (csrc/gemm_sm100.cu): LN(x) W^T + b = rstd (x W'^T) - rstd mean colsum(W') + (b + W beta)
        # with W' = bf16(gamma * W) - in "high" precision the [hi | lo] pair of gamma * W, colsum that of hi + lo;
        # composed in fp64.  The GEMM then runs on the bf16 copy of x its producer left (hi / lo pair in "high") and the
        # stand-alone LayerNorm pass disappears.
        def fold(lin, norm, name):
            W = lin.weight.detach().to(device).double()
            g = norm.weight.detach().to(device).double()
            beta = norm.bias.detach().to(device).double()
            wf = self._gemm_weight((W * g[None, :]).float(), device)
            pk[name + ".wf"] = wf
            K = W.shape[1]
            pk[name + ".sf"] = (wf[:, :K].double() + (wf[:, K:].double() if wf.shape[1] > K else 0.0)).sum(1).float().contiguous()
            pk[name + ".bf"] = (lin.bias.detach().to(device).double() + W @ beta).float().contiguous()
        # head-major row order of the temporal qkv tensors for the fused qkv-GEMM + temporal-attention kernel
        # (csrc/gemm_tattn_sm100.cu, bf16 precision): one 192-row block per head = [q_h | k_h | v_h]
        hp = torch.arange(2304, device=device).view(3, 12, 64).permute(1, 0, 2).reshape(-1)
        for i, blk in enumerate(self.blocks):
            p = f"blocks.{i}."
            if self.attention_type == 'divided_space_time':
                fold(blk.temporal_attn.qkv, blk.temporal_norm1, p + "temporal_attn.qkv")
                if self.precision == "bf16":
                    for sfx in ("wf", "sf", "bf"):
                        pk[p + "temporal_attn.qkv." + sfx + "p"] = pk[p + "temporal_attn.qkv." + sfx][hp].contiguous()
            fold(blk.attn.qkv, blk.norm1, p + "attn.qkv")
            fold(blk.mlp.fc1, blk.norm2, p + "mlp.fc1")
        P = int(round(math.sqrt(self.pos_embed.shape[1] - 1)))
        cfg = _Config(self.depth, self.time_embed.shape[1] if self.attention_type != 'space_only' else 1, P,
                      _ATTN[self.attention_type], _PRECISION[self.precision])
        handle = ctypes.c_void_p()
        _lib.check(L.dvst_create(ctypes.byref(cfg), ctypes.byref(handle)))
        for name, t in pk.items():
            _lib.check(L.dvst_set_weight(handle, name.encode(), _lib.ptr(t)))
        _lib.check(L.dvst_finalize_weights(handle))
        self._handle, self._packed, self._pack_key = handle, pk, key

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.lib().dvst_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------------
    def forward_features(self, x, get_all=False, get_attn=False):
        """[B,3,T,H,W] -> [B,768] (or [B,1+N*T,768] if get_all) - models/timesformer.py:285-345.  Under torch.no_grad()
        (every scoring / feature-extraction caller of the reference) this is the plain kernel path; when a gradient is
        being recorded (a student being trained, train_ssl.py:483-563) the CLS-feature form goes through
        autograd.backbone_features: same kernels forward, PyTorch recomputation backward."""
        if torch.is_grad_enabled() and not get_all and not get_attn and x.is_cuda and \
                (x.requires_grad or any(p.requires_grad for p in self._backbone_params())):
            from ..autograd import backbone_features
            return backbone_features(self, x)
        return self._forward_features_cuda(x, get_all, get_attn)

    def _forward_features_cuda(self, x, get_all=False, get_attn=False):
        if x.dim() != 5 or x.shape[1] != 3:
            raise ValueError(f"expected input [B,3,T,H,W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise _lib.DvstError("dvst_b200 runs on CUDA (sm_100a) only: move the input to a B200 - there is no "
                                 "CPU fallback")
        B, _, T, H, W = x.shape
        with torch.cuda.device(x.device):
            self._ensure_packed(x.device)
            L = _lib.lib()
            xin = x.detach().to(torch.float32).contiguous()
            need = L.dvst_workspace_bytes(self._handle, B, T, H, W)
            if need < 0:
                raise _lib.DvstError(f"bad input shape {tuple(x.shape)}")
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != x.device:
                self._workspace = None
                self._workspace = torch.empty(need, dtype=torch.uint8, device=x.device)
            N = (H // 16) * (W // 16)
            if get_attn:
                # last block's spatial attention map [B*T, 12, 1+N, 1+N] (models/timesformer.py:325-331,155-157)
                if self.attention_type != 'divided_space_time':
                    raise ValueError("get_attn is defined for divided_space_time only (the reference's other block "
                                     "types ignore return_attn, models/timesformer.py:120-123)")
                attn = torch.empty(B * T, 12, N + 1, N + 1, dtype=torch.float32, device=x.device)
                _lib.check(L.dvst_forward_attention_map(self._handle, _lib.ptr(xin), B, T, H, W, _lib.ptr(attn),
                                                        _lib.ptr(self._workspace), self._workspace.numel(),
                                                        _lib.cur_stream()))
                return attn
            if get_all:
                rows = 1 + (N if self.attention_type == 'space_only' else N * T)
                out = torch.empty(B, rows, self.embed_dim, dtype=torch.float32, device=x.device)
            else:
                out = torch.empty(B, self.embed_dim, dtype=torch.float32, device=x.device)
            _lib.check(L.dvst_forward_features(self._handle, _lib.ptr(xin), B, T, H, W, _lib.ptr(out),
                                               1 if get_all else 0, _lib.ptr(self._workspace),
                                               self._workspace.numel(), _lib.cur_stream()))
        return out

    def forward(self, x, use_head=False):
        x = self.forward_features(x)
        if use_head:
            x = self.head(x)
        return x

    def get_intermediate_layers(self, x, n=1):
        return [self.forward_features(x, get_all=True), ]


default_cfgs = {'vit_base_patch16_224': {
    'url': "https://dl.fbaipublicfiles.com/dino/dino_vitbase16_pretrain/dino_vitbase16_pretrain.pth",
    'num_classes': 1000, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9, 'interpolation': 'bicubic',
    'mean': (0.485, 0.456, 0.406), 'std': (0.229, 0.224, 0.225), 'first_conv': 'patch_embed.proj',
    'classifier': 'head'}}


def _load_pretrained_file(model, path, attention_type):
    """Local-file subset of models/helpers.py:load_pretrained (:165-200): clone spatial attention / norm1 into the
    temporal branch when the checkpoint has none.  URL download is not available offline."""
    sd = torch.load(path, map_location="cpu")
    sd = sd.get("model_state", sd.get("model", sd))
    sd = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in sd.items()}
    if attention_type == 'divided_space_time':
        new = dict(sd)
        for k, v in sd.items():
            if 'blocks' in k and 'attn' in k and 'temporal' not in k:
                nk = k.replace('attn', 'temporal_attn')
                new.setdefault(nk, v)
            if 'blocks' in k and 'norm1' in k and 'temporal' not in k:
                nk = k.replace('norm1', 'temporal_norm1')
                new.setdefault(nk, v)
        sd = new
    return model.load_state_dict(sd, strict=False)


def _build(cfg_like, **kwargs):
    return VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_rate=0., attn_drop_rate=0.,
                             drop_path_rate=0.1, **cfg_like, **kwargs)


def get_vit_base_patch16_224(cfg, no_head=False, **kwargs):
    """Same factory as models/timesformer.py:592-609 (cfg: any object exposing DATA.TRAIN_CROP_SIZE, DATA.NUM_FRAMES,
    MODEL.NUM_CLASSES, TIMESFORMER.ATTENTION_TYPE, TIMESFORMER.PRETRAINED_MODEL)."""
    patch_size = 16
    vit = _build(dict(img_size=cfg.DATA.TRAIN_CROP_SIZE, num_classes=cfg.MODEL.NUM_CLASSES,
                      num_frames=cfg.DATA.NUM_FRAMES, attention_type=cfg.TIMESFORMER.ATTENTION_TYPE), **kwargs)
    vit.attention_type = cfg.TIMESFORMER.ATTENTION_TYPE
    vit.default_cfg = default_cfgs['vit_base_patch16_224']
    vit.num_patches = (cfg.DATA.TRAIN_CROP_SIZE // patch_size) * (cfg.DATA.TRAIN_CROP_SIZE // patch_size)
    pretrained_model = cfg.TIMESFORMER.PRETRAINED_MODEL
    if pretrained_model:
        _load_pretrained_file(vit, pretrained_model, vit.attention_type)
    if no_head:
        vit.head = None
    return vit


class vit_base_patch16_224(nn.Module):
    """models/timesformer.py:379-401.  The reference always calls load_pretrained (URL download when
    PRETRAINED_MODEL is empty); offline, an empty path leaves the random init in place."""

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.pretrained = True
        self.model = _build(dict(img_size=cfg.DATA.TRAIN_CROP_SIZE, num_classes=cfg.MODEL.NUM_CLASSES,
                                 num_frames=cfg.DATA.NUM_FRAMES, attention_type=cfg.TIMESFORMER.ATTENTION_TYPE),
                            **kwargs)
        self.attention_type = cfg.TIMESFORMER.ATTENTION_TYPE
        self.model.default_cfg = default_cfgs['vit_base_patch16_224']
        self.num_patches = (cfg.DATA.TRAIN_CROP_SIZE // 16) * (cfg.DATA.TRAIN_CROP_SIZE // 16)
        if cfg.TIMESFORMER.PRETRAINED_MODEL:
            _load_pretrained_file(self.model, cfg.TIMESFORMER.PRETRAINED_MODEL, self.attention_type)

    def forward(self, x):
        return self.model(x)


class TimeSformer(nn.Module):
    """models/timesformer.py:404-426: forward(x) returns the backbone FEATURES (use_head defaults to False)."""

    def __init__(self, img_size=224, patch_size=16, num_classes=400, num_frames=8, attention_type='divided_space_time',
                 pretrained_model='', **kwargs):
        super().__init__()
        if patch_size != 16:
            raise ValueError("patch_size must be 16")
        self.pretrained = True
        self.model = _build(dict(img_size=img_size, num_classes=num_classes, num_frames=num_frames,
                                 attention_type=attention_type), **kwargs)
        self.attention_type = attention_type
        self.model.default_cfg = default_cfgs['vit_base_patch16_224']
        self.num_patches = (img_size // patch_size) * (img_size // patch_size)
        if pretrained_model:
            _load_pretrained_file(self.model, pretrained_model, attention_type)

    def forward(self, x):
        return self.model(x)
