"""Forward + center update of the training DINOLoss (train_ssl.py:602-691, two_token=False) on the CUDA kernels.

Same constructor / forward(student_output, teacher_output, epoch) / update_center / `center` buffer as the reference.
The cross-entropy terms run through dvst_pair_loss_indexed (one launch for all (teacher view, student view) terms),
the batch-center column sum and the EMA are CUDA kernels, and the all-reduce in between is NCCL via torch.distributed
exactly where the reference has it (train_ssl.py:687).  The value is computed by the kernels; its gradient w.r.t. the
student output by dvst_b200/autograd.py (PyTorch).  The SSL training loop itself is out of scope.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib


class DINOLoss(nn.Module):
    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs,
                 student_temp=0.1, center_momentum=0.9, global_crops=2, two_token=False):
        super().__init__()
        if two_token:
            raise ValueError("two_token=True (MODEL.TWO_TOKEN) is outside the accelerated path")
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.n_crops = ncrops
        self.global_crops = global_crops
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp))

    def forward(self, student_output, teacher_output, epoch):
        """train_ssl.py:626-671.  The value comes from the fused pair-loss kernel; when the student output carries a graph
        the gradient w.r.t. it is supplied by autograd._TrainLoss (PyTorch), so `loss.backward()` works as in the
        reference's training step (train_ssl.py:524-540)."""
        if torch.is_grad_enabled() and student_output.requires_grad:
            from .autograd import _TrainLoss
            return _TrainLoss.apply(self, student_output, teacher_output, epoch)
        return self._forward_cuda(student_output, teacher_output, epoch)

    @torch.no_grad()
    def _forward_cuda(self, student_output, teacher_output, epoch):
        if not student_output.is_cuda:
            raise _lib.DvstError("dvst_b200 runs on CUDA (sm_100a) only - there is no CPU fallback")
        L = _lib.lib()
        s = student_output.detach().float().contiguous()
        t = teacher_output.detach().float().contiguous()
        dim = s.shape[-1]
        B = s.shape[0] // self.n_crops
        if t.shape[0] != B * self.global_crops:
            raise ValueError("teacher_output must hold global_crops * B rows")
        temp = float(self.teacher_temp_schedule[epoch])
        s_idx, t_idx = [], []
        for iq in range(self.global_crops):
            for v in range(self.n_crops):
                if v == iq:
                    continue  # student and teacher on the same view
                s_idx += list(range(v * B, (v + 1) * B))
                t_idx += list(range(iq * B, (iq + 1) * B))
        n_terms = len(s_idx) // B
        si = torch.tensor(s_idx, dtype=torch.int32, device=s.device)
        ti = torch.tensor(t_idx, dtype=torch.int32, device=s.device)
        losses = torch.empty(len(s_idx), dtype=torch.float32, device=s.device)
        center = self.center.float().contiguous()
        with torch.cuda.device(s.device):
            _lib.check(L.dvst_pair_loss_indexed(_lib.ptr(s), s.stride(0), _lib.ptr(si), _lib.ptr(t), t.stride(0),
                                                _lib.ptr(ti), _lib.ptr(center), temp, float(self.student_temp), dim,
                                                len(s_idx), _lib.ptr(losses), _lib.cur_stream()))
        # sum over terms of the per-term batch mean, divided by the number of terms (train_ssl.py:665-668)
        total_loss = losses.reshape(n_terms, B).mean(dim=1).sum() / n_terms
        self.update_center(t)
        return total_loss

    @torch.no_grad()
    def update_center(self, teacher_output):
        """train_ssl.py:673-691: column sum -> all_reduce(SUM) -> / (B * world) -> EMA."""
        L = _lib.lib()
        t = teacher_output.detach().float().contiguous()
        dim = t.shape[-1]
        batch_center = torch.empty(1, dim, dtype=torch.float32, device=t.device)
        with torch.cuda.device(t.device):
            _lib.check(L.dvst_colsum(_lib.ptr(t), t.stride(0), _lib.ptr(batch_center), t.shape[0], dim,
                                     _lib.cur_stream()))
            world = 1
            if dist.is_available() and dist.is_initialized():
                dist.all_reduce(batch_center)
                world = dist.get_world_size()
            center = self.center.float().contiguous()
            _lib.check(L.dvst_center_ema(_lib.ptr(center), _lib.ptr(batch_center), float(t.shape[0] * world),
                                         float(self.center_momentum), dim, _lib.cur_stream()))
        self.center = center
