"""Autograd glue for the training-mode forward ("forward on the kernels, backward still
PyTorch"): `train_ssl.py:483-563` back-propagates the DINO loss through the student's head and backbone.

The FORWARD of every piece below is the product path (libdvst.so: tcgen05 GEMMs, fused attention, fused loss kernels).
The BACKWARD is PyTorch: each `torch.autograd.Function` keeps its inputs, and when a gradient arrives it re-evaluates the
same function with differentiable torch operations on the device (activation recomputation - nothing but the inputs is
kept alive between forward and backward) and lets torch.autograd differentiate that.  The gradient is therefore the exact
fp32 gradient of the reference formulation at the current parameters; the forward value differs from it by the bf16
operand rounding of the kernels (the usual mixed-precision straight-through arrangement).

  * `backbone_features(module, x)`      - VisionTransformer.forward_features (models/timesformer.py:285-345), CLS features
  * `head_logits(head, feats)`          - DINOHead.forward (vision_transformer.py:278-282)
  * `dino_train_loss(...)`              - training DINOLoss.forward (train_ssl.py:626-671), gradient w.r.t. the student

Stochastic depth (`drop_path`, models/vit_utils.py:140-161) is not part of the kernels' forward: the student runs as if
drop_path_rate = 0.  Nothing here is used by the scoring path (which runs under torch.no_grad()).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _nearest_index(dst_len, src_len, device):
    """Index map of F.interpolate(mode='nearest') (models/timesformer.py:292-303,314-318), computed in fp32 like ATen."""
    scale = np.float32(src_len) / np.float32(dst_len)
    idx = np.floor(np.arange(dst_len, dtype=np.float32) * scale).astype(np.int64)
    return torch.from_numpy(np.minimum(idx, src_len - 1)).to(device)


def _attention(mod, x):
    """Attention.forward (models/timesformer.py:68-87) on x [..., S, 768]."""
    lead, S = x.shape[:-2], x.shape[-2]
    qkv = F.linear(x, mod.qkv.weight, mod.qkv.bias).reshape(*lead, S, 3, mod.num_heads, -1)
    q, k, v = (qkv[..., i, :, :].transpose(-2, -3) for i in range(3))          # [..., heads, S, 64]
    a = F.scaled_dot_product_attention(q, k, v)                                 # scale = 64 ** -0.5
    return F.linear(a.transpose(-2, -3).reshape(*lead, S, -1), mod.proj.weight, mod.proj.bias)


def _ln(mod, x):
    return F.layer_norm(x, (x.shape[-1],), mod.weight, mod.bias, mod.eps)


def _mlp(mod, x):
    return F.linear(F.gelu(F.linear(x, mod.fc1.weight, mod.fc1.bias)), mod.fc2.weight, mod.fc2.bias)


def torch_forward_features(m, x):
    """Differentiable torch statement of VisionTransformer.forward_features(x)[CLS] for the backward pass, on the
    resident [B, N, T, D] token layout + separate CLS rows (same arithmetic as models/timesformer.py:285-345)."""
    B, _, T, H, W = x.shape
    gh, gw = H // 16, W // 16
    N = gh * gw
    w = m.patch_embed.proj.weight.reshape(768, -1)
    p = x.reshape(B, 3, T, gh, 16, gw, 16).permute(0, 2, 3, 5, 1, 4, 6).reshape(B, T, N, 768)
    tok = F.linear(p, w, m.patch_embed.proj.bias)                               # [B, T, N, D]
    P = int(round(math.sqrt(m.pos_embed.shape[1] - 1)))
    pos = m.pos_embed[0, 1:]
    if N != P * P:
        rows, cols = _nearest_index(gh, P, x.device), _nearest_index(gw, P, x.device)
        pos = pos.reshape(P, P, -1)[rows][:, cols].reshape(N, -1)
    tok = tok + pos[None, None]
    cls = (m.cls_token[0, 0] + m.pos_embed[0, 0])[None].expand(B, -1)
    att = m.attention_type
    if att == 'space_only':
        seq = torch.cat([cls[:, None, None].expand(B, T, 1, 768), tok], 2)       # [B, T, 1 + N, D]
        for blk in m.blocks:
            seq = seq + _attention(blk.attn, _ln(blk.norm1, seq))
            seq = seq + _mlp(blk.mlp, _ln(blk.norm2, seq))
        return _ln(m.norm, seq.mean(1))[:, 0]
    te = m.time_embed[0]
    if T != te.shape[0]:
        te = te[_nearest_index(T, te.shape[0], x.device)]
    tok = (tok + te[None, :, None]).transpose(1, 2)                             # [B, N, T, D]
    if att == 'joint_space_time':
        seq = torch.cat([cls[:, None], tok.reshape(B, N * T, 768)], 1)
        for blk in m.blocks:
            seq = seq + _attention(blk.attn, _ln(blk.norm1, seq))
            seq = seq + _mlp(blk.mlp, _ln(blk.norm2, seq))
        return _ln(m.norm, seq[:, 0])
    for blk in m.blocks:
        r = _attention(blk.temporal_attn, _ln(blk.temporal_norm1, tok))        # over T for every (b, n)
        tok = tok + F.linear(r, blk.temporal_fc.weight, blk.temporal_fc.bias)
        seq = torch.cat([cls[:, None, None].expand(B, T, 1, 768), tok.transpose(1, 2)], 2)   # [B, T, 1 + N, D]
        r = _attention(blk.attn, _ln(blk.norm1, seq))
        cls = cls + r[:, :, 0].mean(1)
        tok = tok + r[:, :, 1:].transpose(1, 2)
        full = torch.cat([cls[:, None], tok.reshape(B, N * T, 768)], 1)
        full = full + _mlp(blk.mlp, _ln(blk.norm2, full))
        cls, tok = full[:, 0], full[:, 1:].reshape(B, N, T, 768)
    return _ln(m.norm, cls)


def torch_head_logits(head, x):
    """DINOHead.forward (vision_transformer.py:278-282) with the weight-normed last layer g * v / ||v||."""
    x = head.mlp(x)
    x = F.normalize(x, dim=-1, p=2)
    ll = head.last_layer
    w = ll.weight_g * ll.weight_v / ll.weight_v.norm(dim=1, keepdim=True)
    return F.linear(x, w)


class _Recompute(torch.autograd.Function):
    """forward: `cuda_fn(x)` (the kernels, no graph); backward: torch.autograd through `torch_fn(x)` recomputed on the
    spot.  `params` are passed as inputs so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, cuda_fn, torch_fn, x, *params):
        ctx.torch_fn, ctx.params = torch_fn, params
        ctx.save_for_backward(x)
        with torch.no_grad():
            return cuda_fn(x)

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        need = [i for i, p in enumerate(ctx.params) if p.requires_grad]
        with torch.enable_grad():
            xin = x.detach().requires_grad_(ctx.needs_input_grad[2])
            out = ctx.torch_fn(xin)
            wrt = ([xin] if ctx.needs_input_grad[2] else []) + [ctx.params[i] for i in need]
            grads = torch.autograd.grad(out, wrt, grad_out.to(out.dtype), allow_unused=True)
        gx = grads[0] if ctx.needs_input_grad[2] else None
        gp = [None] * len(ctx.params)
        for i, g in zip(need, grads[1 if ctx.needs_input_grad[2] else 0:]):
            gp[i] = g
        return (None, None, gx, *gp)


def backbone_features(module, x):
    """CLS features [B, 768] of a dvst_b200 VisionTransformer with a PyTorch backward."""
    params = tuple(p for n, p in module.named_parameters() if not n.startswith("head."))
    return _Recompute.apply(lambda t: module._forward_features_cuda(t), lambda t: torch_forward_features(module, t.float()),
                            x, *params)


def head_logits(head, feats):
    params = tuple(head.parameters())
    return _Recompute.apply(lambda t: head._forward_cuda(t), lambda t: torch_head_logits(head, t.float()), feats, *params)


class _TrainLoss(torch.autograd.Function):
    """Training DINOLoss value from the fused pair-loss kernel; gradient w.r.t. the student logits in torch:
    d/ds_v = (1 / (n_terms B tau_s)) * sum over teacher views iq != v of (softmax(s_v / tau_s) - q_iq)."""

    @staticmethod
    def forward(ctx, loss_module, student_output, teacher_output, epoch):
        ctx.mod, ctx.epoch = loss_module, epoch
        ctx.save_for_backward(student_output, teacher_output, loss_module.center.clone())
        return loss_module._forward_cuda(student_output, teacher_output, epoch)

    @staticmethod
    def backward(ctx, grad_out):
        s, t, center = ctx.saved_tensors
        mod = ctx.mod
        temp = float(mod.teacher_temp_schedule[ctx.epoch])
        B = s.shape[0] // mod.n_crops
        q = F.softmax((t.float() - center.float()) / temp, dim=-1).reshape(mod.global_crops, B, -1)
        ps = F.softmax(s.float() / mod.student_temp, dim=-1).reshape(mod.n_crops, B, -1)
        n_terms = mod.global_crops * mod.n_crops - min(mod.global_crops, mod.n_crops)
        g = torch.zeros_like(ps)
        for v in range(mod.n_crops):
            k = mod.global_crops - (1 if v < mod.global_crops else 0)          # teacher views paired with student view v
            qs = q.sum(0) - (q[v] if v < mod.global_crops else 0)
            g[v] = k * ps[v] - qs
        g = g.reshape_as(s) * (grad_out / (n_terms * B * mod.student_temp))
        return None, g.to(s.dtype), None, None
