"""Drop-in for the reference's `vision_transformer.DINOHead` (vision_transformer.py:248-282).

Same constructor, same state_dict keys (`mlp.{0,2,4}.{weight,bias}`, `last_layer.weight_g`, `last_layer.weight_v` - the
deprecated weight_norm hook layout).  `forward(x)` returns the materialised logits [B, out_dim] like the reference
(needed by MultiCropWrapper / the training loss); the scoring path instead calls `bottleneck(x)` + the fused
`dvst_head_loss` kernel, which never materialises them (see dino_similarity.HeadScorer).
All GEMMs run on the tcgen05 kernel of csrc/gemm_sm100.cu through the C ABI.
"""
import torch
import torch.nn as nn

from . import _lib
from .models.timesformer import trunc_normal_


class _WeightNormLinear(nn.Module):
    """Parameter container reproducing nn.utils.weight_norm(nn.Linear(in, out, bias=False)) key names."""

    def __init__(self, in_features, out_features):
        super().__init__()
        lin = nn.Linear(in_features, out_features, bias=False)
        self.weight_g = nn.Parameter(lin.weight.detach().norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(lin.weight.detach().clone())

    @property
    def weight(self):
        v = self.weight_v
        return self.weight_g * v / v.norm(dim=1, keepdim=True)


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256, precision="bf16"):
        super().__init__()
        if use_bn:
            raise ValueError("use_bn=True is not supported by the fused head (the reference scripts use use_bn=False)")
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim)
        else:
            layers = [nn.Linear(in_dim, hidden_dim), nn.GELU()]
            for _ in range(nlayers - 2):
                layers += [nn.Linear(hidden_dim, hidden_dim), nn.GELU()]
            layers.append(nn.Linear(hidden_dim, bottleneck_dim))
            self.mlp = nn.Sequential(*layers)
        self.apply(self._init_weights)
        self.last_layer = _WeightNormLinear(bottleneck_dim, out_dim)
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False
        self.in_dim, self.out_dim, self.bottleneck_dim = in_dim, out_dim, bottleneck_dim
        for d in (hidden_dim if nlayers > 1 else 256, bottleneck_dim, out_dim):
            if d % 256:
                raise ValueError("hidden_dim, bottleneck_dim and out_dim must be multiples of 256 (GEMM tile width)")
        if in_dim % 64 or bottleneck_dim != 256:
            raise ValueError("in_dim must be a multiple of 64 and bottleneck_dim 256 (fused loss kernel)")
        self.precision = precision
        self._pack_key, self._packed = None, None

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------------------------------------------------------------
    def _linears(self):
        return [self.mlp] if isinstance(self.mlp, nn.Linear) else [m for m in self.mlp if isinstance(m, nn.Linear)]

    def _ensure_packed(self, device):
        ps = list(self.parameters())
        key = (str(device), self.precision, tuple((p.data_ptr(), p._version) for p in ps))
        if key == self._pack_key:
            return self._packed
        split = self.precision == "high"

        def gw(w):
            w = w.detach().to(device=device, dtype=torch.float32)
            if split:
                hi = w.to(torch.bfloat16)
                return torch.cat([hi, (w - hi.float()).to(torch.bfloat16)], dim=1).contiguous()
            return w.to(torch.bfloat16).contiguous()
        pk = {"mlp": [(gw(l.weight), l.bias.detach().to(device=device, dtype=torch.float32).contiguous())
                      for l in self._linears()]}
        # weight_norm recomputes g * v / ||v|| on every forward in the reference; the head is frozen here -> once
        w_last = self.last_layer.weight.detach().to(device=device, dtype=torch.float32)
        pk["last"] = gw(w_last)
        pk["last_bf16"] = w_last.to(torch.bfloat16).contiguous()  # operand of the fused loss kernel
        self._pack_key, self._packed = key, pk
        return pk

    def bottleneck(self, x):
        """mlp(x) followed by F.normalize (vision_transformer.py:279-280) -> (bf16 operand [B, 256*k], fp32 copy)."""
        if not x.is_cuda:
            raise _lib.DvstError("dvst_b200 runs on CUDA (sm_100a) only - there is no CPU fallback")
        L = _lib.lib()
        pk = self._ensure_packed(x.device)
        split = 1 if self.precision == "high" else 0
        k = 2 if split else 1
        x = x.detach().to(torch.float32).contiguous()
        B = x.shape[0]
        st = _lib.cur_stream()
        a = torch.empty(B, self.in_dim * k, dtype=torch.bfloat16, device=x.device)
        _lib.check(L.dvst_cast_bf16(_lib.ptr(x), x.stride(0), _lib.ptr(a), a.stride(0), split, self.in_dim, B,
                                    self.in_dim, st))
        n = len(pk["mlp"])
        for i, (w, b) in enumerate(pk["mlp"]):
            N, K = w.shape[0], w.shape[1] // k
            last = i == n - 1
            if last:
                out = torch.empty(B, N, dtype=torch.float32, device=x.device)
                _lib.check(L.dvst_gemm_bf16(_lib.ptr(a), a.stride(0), _lib.ptr(w), w.stride(0), _lib.ptr(b), B, N, K,
                                            split, None, 0, _lib.ptr(out), N, 1, 0, 0, st))
            else:
                out = torch.empty(B, N * k, dtype=torch.bfloat16, device=x.device)
                _lib.check(L.dvst_gemm_bf16(_lib.ptr(a), a.stride(0), _lib.ptr(w), w.stride(0), _lib.ptr(b), B, N, K,
                                            split, None, 0, _lib.ptr(out), N * k, 2 if split else 0, N, 1, st))
            a = out
        z = torch.empty(B, self.bottleneck_dim * k, dtype=torch.bfloat16, device=x.device)
        zf = torch.empty(B, self.bottleneck_dim, dtype=torch.float32, device=x.device)
        _lib.check(L.dvst_l2norm_rows(_lib.ptr(a), a.stride(0), _lib.ptr(z), z.stride(0), split, self.bottleneck_dim,
                                      _lib.ptr(zf), B, self.bottleneck_dim, st))
        return z, zf

    def forward(self, x):
        """Logits [B, out_dim] (vision_transformer.py:278-282).  With gradients recorded: kernels forward, PyTorch
        recomputation backward (autograd.head_logits)."""
        if torch.is_grad_enabled() and x.is_cuda and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import head_logits
            return head_logits(self, x)
        return self._forward_cuda(x)

    def _forward_cuda(self, x):
        z, _ = self.bottleneck(x)
        pk = self._packed
        L = _lib.lib()
        split = 1 if self.precision == "high" else 0
        B = x.shape[0]
        out = torch.empty(B, self.out_dim, dtype=torch.float32, device=x.device)
        w = pk["last"]
        _lib.check(L.dvst_gemm_bf16(_lib.ptr(z), z.stride(0), _lib.ptr(w), w.stride(0), None, B, self.out_dim,
                                    self.bottleneck_dim, split, None, 0, _lib.ptr(out), self.out_dim, 1, 0, 0,
                                    _lib.cur_stream()))
        return out
