"""Drop-in for the reference's dino_similarity.py: per-frame DINO-loss scoring of a video for adaptive frame selection.

  * `DINOLoss`        - dino_similarity.py:120-135 (same constructor / forward / `center` buffer).
  * `export_loss`     - dino_similarity.py:97-117 (same JSON read-modify-write format).
  * `score_video`     - dino_similarity.py:66-93 for ONE video, restructured for the GPU: frames are uploaded once
                        ([F,3,224,224], not the reference's [2F,3,G,224,224] padded views), windows are built on the
                        device by index (dvst_forward_windows), identical windows are scored once (only F-2 student and
                        max(F-G+1,1) teacher windows are distinct) and all F losses come back in a single D2H copy
                        instead of one `.item()` sync per frame.  Results are unchanged: every frame's loss is the loss
                        of exactly the (local, global) window pair the reference builds for it.
  * `score_videos_device` - the same for a batch of videos in one backbone pass per window length (bit-identical
                        losses; small videos no longer run the GPU at batch 1).
  * `dino_similarity` - the driver loop over a dataset of videos.

Mode R (reference-exact): loss over the 768-d backbone features, tau_t = 0.02, tau_s = 0.3, zero center.
Mode H (north star): `head=DINOHead(768, 65536)`; features -> bottleneck -> fused last-layer GEMM + softmax + CE.
"""
import json
import math
import os

import torch
import torch.nn as nn

from . import _lib


def _check_indices(idx, n_rows, device, validate=None):
    """Row-gather indices of the C ABI are int32 device arrays (include/dvst.h): anything else would be read as garbage.
    The range check costs a device sync, so it runs only when DVST_CHECK_INDICES=1 (tests) or validate=True."""
    if idx is None:
        raise ValueError("s_idx and t_idx must be given together")
    if idx.dtype != torch.int32 or idx.device != device or idx.dim() != 1:
        raise ValueError(f"row indices must be 1-D int32 tensors on {device} (got {idx.dtype} on {idx.device})")
    idx = idx.contiguous()
    if validate is None:
        validate = os.environ.get("DVST_CHECK_INDICES") == "1"
    if validate and idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= n_rows):
        raise IndexError(f"row index out of range [0, {n_rows})")
    return idx


class DINOLoss(nn.Module):
    """calculates the DINO loss for two output tensors and returns a single value (dino_similarity.py:120-135)."""

    def __init__(self, out_dim, teacher_temp, student_temp=0.1):
        super().__init__()
        self.student_temp = student_temp
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.teacher_temp = teacher_temp

    def pair_losses(self, student_output, teacher_output, s_idx=None, t_idx=None):
        """Per-pair losses [P] for rows of student_output / teacher_output ([P, out_dim] fp32 CUDA tensors); optional
        int32 row-gather indices pair student row s_idx[p] with teacher row t_idx[p]."""
        if not student_output.is_cuda:
            raise _lib.DvstError("dvst_b200 runs on CUDA (sm_100a) only - there is no CPU fallback")
        s = student_output.detach().to(torch.float32)
        t = teacher_output.detach().to(torch.float32)
        s = s.reshape(-1, s.shape[-1]).contiguous()
        t = t.reshape(-1, t.shape[-1]).contiguous()
        dim = s.shape[-1]
        if t.shape[-1] != dim or self.center.shape[-1] != dim:
            raise ValueError("student / teacher / center dims differ")
        P = s.shape[0] if s_idx is None else s_idx.numel()
        if s_idx is None and t.shape[0] != P:
            raise ValueError("student and teacher must hold the same number of rows")
        if s_idx is not None:
            s_idx, t_idx = _check_indices(s_idx, s.shape[0], s.device), _check_indices(t_idx, t.shape[0], s.device)
            if t_idx.numel() != P:
                raise ValueError("s_idx and t_idx must have the same length")
        out = torch.empty(P, dtype=torch.float32, device=s.device)
        center = self.center.to(device=s.device, dtype=torch.float32).contiguous()
        L = _lib.lib()
        with torch.cuda.device(s.device):
            if s_idx is None:
                _lib.check(L.dvst_pair_loss(_lib.ptr(s), s.stride(0), _lib.ptr(t), t.stride(0), _lib.ptr(center),
                                            float(self.teacher_temp), float(self.student_temp), dim, P, _lib.ptr(out),
                                            _lib.cur_stream()))
            else:
                _lib.check(L.dvst_pair_loss_indexed(_lib.ptr(s), s.stride(0), _lib.ptr(s_idx), _lib.ptr(t),
                                                    t.stride(0), _lib.ptr(t_idx), _lib.ptr(center),
                                                    float(self.teacher_temp), float(self.student_temp), dim, P,
                                                    _lib.ptr(out), _lib.cur_stream()))
        return out

    def forward(self, student_output, teacher_output):
        # cross-entropy between softmax outputs for the teacher and student; `.mean()` over pairs like the reference
        return self.pair_losses(student_output, teacher_output).mean()


def export_loss(loss_list, video_path, file_path):
    """exporting the loss values as a JSON file - same format and merge semantics as dino_similarity.py:97-117."""
    video_name = os.path.basename(video_path)
    video_name_without_extension, _ = os.path.splitext(video_name)
    video_dict = {video_name_without_extension: loss_list}
    if os.path.exists(file_path):
        with open(file_path, 'r') as file:
            data = json.load(file)
        merged_dict = data.copy()
        merged_dict.update(video_dict)
        with open(file_path, 'w') as file:
            json.dump(merged_dict, file)
    else:
        with open(file_path, 'w') as file:
            json.dump(video_dict, file)


def export_losses(losses_by_video, file_path):
    """export_loss for several videos with ONE read-modify-write of the JSON (the reference rewrites the whole file
    once per video, dino_similarity.py:97-117 - quadratic in the number of videos).  `losses_by_video` is an ordered
    mapping video_path -> loss list; the resulting file is identical to calling export_loss for each entry in order."""
    video_dict = {os.path.splitext(os.path.basename(path))[0]: loss for path, loss in losses_by_video.items()}
    merged = {}
    if os.path.exists(file_path):
        with open(file_path, 'r') as file:
            merged = json.load(file).copy()
    merged.update(video_dict)
    with open(file_path, 'w') as file:
        json.dump(merged, file)


def preprocess_frames(video_uint8, sampling_rate=4, mean=(0.45, 0.45, 0.45), std=(0.225, 0.225, 0.225), crop_size=224):
    """Loader front half (datasets_custom/dino_loss_loader.py:80-92): raw decoded video uint8 [n,H,W,3] (a CUDA tensor:
    upload uint8, not fp32) -> every `sampling_rate`-th frame, /255, mean/std normalised (data_utils.py:308-325),
    T H W C -> T C H W, centre crop (transform.py:206-249, spatial_idx=1).  One fused kernel (csrc/selection.cu:
    preprocess_u8_kernel, bit-identical to the reference's three fp32 operations) instead of four full-tensor passes.
    There is no CPU / PyTorch path: other inputs raise."""
    if not (video_uint8.is_cuda and video_uint8.dtype == torch.uint8 and video_uint8.dim() == 4
            and video_uint8.shape[3] == 3):
        raise _lib.DvstError("preprocess_frames: expected a CUDA uint8 tensor [n,H,W,3] (decoded frames); dvst_b200 runs "
                             "on CUDA (sm_100a) only - there is no CPU fallback")
    import ctypes
    v = video_uint8.contiguous()
    n, h, w = v.shape[0], v.shape[1], v.shape[2]
    out = torch.empty(-(-n // sampling_rate), 3, crop_size, crop_size, dtype=torch.float32, device=v.device)
    m3 = (ctypes.c_float * 3)(*[float(x) for x in mean])
    s3 = (ctypes.c_float * 3)(*[float(x) for x in std])
    with torch.cuda.device(v.device):
        _lib.check(_lib.lib().dvst_preprocess_u8(_lib.ptr(v), n, h, w, sampling_rate, m3, s3, crop_size,
                                                 _lib.ptr(out), _lib.cur_stream()))
    return out


def window_starts(F, local_clip_size=3, global_clip_size=30):
    """Window rule of get_views_of_video_same_size (datasets_custom/dino_loss_loader.py:126-171) as index arithmetic:
    returns (G, local_starts[F], global_starts[F]); frame i is scored on frames [j, j+local) vs [l, l+G)."""
    loc, glob = int(local_clip_size / 2), int(global_clip_size / 2)
    G = min(global_clip_size, F)
    ls, gs = [], []
    for i in range(F):
        j, k, l, m = i - loc, i + loc + 1, i - glob, i + glob
        if j < 0:
            j, k = 0, local_clip_size
        if k >= F:
            k, j = F, F - local_clip_size
        if l < 0:
            l, m = 0, G
        if m >= F:
            m, l = F, F - G
        ls.append(j)
        gs.append(l)
    return G, ls, gs


def short_video_clips(F, local_clip_size=3):
    """(student frame indices, teacher frame indices) for a video with F < local_clip_size sampled frames.  The reference
    does not treat it as malformed: get_views_of_video_same_size clamps G = F, its local slice frames[F-local : F] has a
    negative start (python wrap-around -> max(2F - local, 0)) and is written into the zero [3,G,H,W] tensor by broadcast
    (dino_loss_loader.py:126-161), so with local = 3 a 2-frame video is scored on student [f1, f1] vs teacher [f0, f1]
    and a 1-frame video on [f0] vs [f0]; every frame gets that same pair.  Slices the reference cannot broadcast raise."""
    start = max(2 * F - local_clip_size, 0)
    n = F - start
    if n == 1:
        s_idx = [start] * F
    elif n == F:
        s_idx = list(range(F))
    else:
        raise ValueError(f"the reference cannot build views for {F} frames with local_clip_size {local_clip_size} "
                         f"(broadcast of {n} frames into {F})")
    return s_idx, list(range(F))


class HeadScorer:
    """Mode-H loss: DINOHead bottleneck + fused last layer / softmax / cross-entropy (never materialises logits)."""

    def __init__(self, head, teacher_temp=0.04, student_temp=0.1, center=None):
        self.head, self.teacher_temp, self.student_temp, self.center = head, teacher_temp, student_temp, center

    def pair_losses(self, feat_s, feat_t, s_idx, t_idx):
        L = _lib.lib()
        zs, _ = self.head.bottleneck(feat_s)
        zt = zs if feat_t is feat_s else self.head.bottleneck(feat_t)[0]
        # gather rows per frame (tiny: [F, 256] bf16)
        s_idx = _check_indices(s_idx, zs.shape[0], zs.device)
        t_idx = _check_indices(t_idx, zt.shape[0], zs.device)
        zs = zs[:, :256].index_select(0, s_idx.long()).contiguous()   # index_select range-checks on the device
        zt = zt[:, :256].index_select(0, t_idx.long()).contiguous()
        P = zs.shape[0]
        w = self.head._packed["last_bf16"]
        loss = torch.empty(P, dtype=torch.float32, device=zs.device)
        nb = L.dvst_head_loss_scratch_bytes(P)
        scratch = torch.empty(nb, dtype=torch.uint8, device=zs.device)
        c = None if self.center is None else self.center.to(device=zs.device, dtype=torch.float32).reshape(-1).contiguous()
        _lib.check(L.dvst_head_loss(_lib.ptr(zs), _lib.ptr(zt), zs.stride(0), _lib.ptr(w), _lib.ptr(c),
                                    float(self.teacher_temp), float(self.student_temp), P, self.head.out_dim,
                                    _lib.ptr(loss), _lib.ptr(scratch), nb, _lib.cur_stream()))
        return loss


def _forward_windows(model, frames, starts, T, max_clips, out=None, ws="_workspace"):
    """Features [n, 768] of the T-frame windows of `frames` [F,3,H,W] starting at `starts` (python ints, or an int32
    CUDA tensor as written by dvst_plan_windows - then nothing is copied from the host)."""
    L = _lib.lib()
    dev = frames.device
    model._ensure_packed(dev)
    F, _, H, W = frames.shape
    if torch.is_tensor(starts):
        if starts.dtype != torch.int32 or starts.device != dev:
            raise ValueError("window starts must be an int32 tensor on the frames' device")
        starts_t = starts.contiguous()
    else:
        starts_t = torch.tensor(starts, dtype=torch.int32, device=dev)
    n = starts_t.numel()
    if out is None:
        out = torch.empty(n, model.embed_dim, dtype=torch.float32, device=dev)
    for c0 in range(0, n, max_clips):
        nb = min(max_clips, n - c0)
        need = L.dvst_workspace_bytes(model._handle, nb, T, H, W)
        buf = getattr(model, ws, None)
        if buf is None or buf.numel() < need or buf.device != dev:
            setattr(model, ws, None)
            buf = torch.empty(need, dtype=torch.uint8, device=dev)
            setattr(model, ws, buf)
        _lib.check(L.dvst_forward_windows(model._handle, _lib.ptr(frames), F, _lib.ptr(starts_t[c0:]), nb, T, H, W,
                                          _lib.ptr(out[c0:]), _lib.ptr(buf), buf.numel(), _lib.cur_stream()))
    return out


def forward_windows_u8(model, video_uint8, starts, T, max_clips=64, sampling_rate=1, mean=(0.45, 0.45, 0.45),
                       std=(0.225, 0.225, 0.225), crop_size=224, out=None, ws="_workspace"):
    """Features [n, 768] of T-frame windows cut straight from DECODED frames (uint8 [n_frames,H,W,3] on the device): the
    loader front half (sub-sampling, / 255, mean / std, centre crop - datasets_custom/dino_loss_loader.py:80-92) is applied
    inside the patch extraction (dvst_forward_windows_u8), bit-identical to preprocess_frames() + _forward_windows() without
    the normalised fp32 frames in between.  `starts` index the sampled frame axis."""
    if not (video_uint8.is_cuda and video_uint8.dtype == torch.uint8 and video_uint8.dim() == 4
            and video_uint8.shape[3] == 3):
        raise _lib.DvstError("forward_windows_u8: expected a CUDA uint8 tensor [n,H,W,3] (decoded frames); dvst_b200 runs on "
                             "CUDA (sm_100a) only - there is no CPU fallback")
    import ctypes
    L = _lib.lib()
    dev = video_uint8.device
    model._ensure_packed(dev)
    v = video_uint8.contiguous()
    n_frames, hf, wf = v.shape[0], v.shape[1], v.shape[2]
    if torch.is_tensor(starts):
        if starts.dtype != torch.int32 or starts.device != dev:
            raise ValueError("window starts must be an int32 tensor on the frames' device")
        starts_t = starts.contiguous()
    else:
        starts_t = torch.tensor(starts, dtype=torch.int32, device=dev)
    n = starts_t.numel()
    if out is None:
        out = torch.empty(n, model.embed_dim, dtype=torch.float32, device=dev)
    m3 = (ctypes.c_float * 3)(*[float(x) for x in mean])
    s3 = (ctypes.c_float * 3)(*[float(x) for x in std])
    with torch.cuda.device(dev):
        for c0 in range(0, n, max_clips):
            nb = min(max_clips, n - c0)
            need = L.dvst_workspace_bytes(model._handle, nb, T, crop_size, crop_size)
            buf = getattr(model, ws, None)
            if buf is None or buf.numel() < need or buf.device != dev:
                setattr(model, ws, None)
                buf = torch.empty(need, dtype=torch.uint8, device=dev)
                setattr(model, ws, buf)
            _lib.check(L.dvst_forward_windows_u8(model._handle, _lib.ptr(v), n_frames, hf, wf, sampling_rate, m3, s3,
                                                 crop_size, _lib.ptr(starts_t[c0:]), nb, T, _lib.ptr(out[c0:]),
                                                 _lib.ptr(buf), buf.numel(), _lib.cur_stream()))
    return out


def plan_windows(lengths, local_clip_size, global_clip_size, device):
    """dvst_plan_windows: the loader's window rule for a batch of videos (concatenated along the frame axis) as device
    index arrays.  Returns (s_starts, t_starts, s_idx, t_idx, groups) with groups = [(G, n_windows, first_row), ...]
    ascending in G.  The lengths travel as kernel parameters: no host-to-device copy, no synchronisation."""
    import ctypes
    n = len(lengths)
    n_s = sum(F - local_clip_size + 1 for F in lengths)
    n_t = sum(F - min(global_clip_size, F) + 1 for F in lengths)
    total = sum(lengths)
    buf = torch.empty(n_s + n_t + 2 * total, dtype=torch.int32, device=device)
    s_starts, t_starts = buf[:n_s], buf[n_s:n_s + n_t]
    s_idx, t_idx = buf[n_s + n_t:n_s + n_t + total], buf[n_s + n_t + total:]
    arr = (ctypes.c_int * n)(*[int(F) for F in lengths])
    o_ns, o_nt, o_ng = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    groups = (ctypes.c_int * (3 * n))()
    with torch.cuda.device(device):
        _lib.check(_lib.lib().dvst_plan_windows(arr, n, int(local_clip_size), int(global_clip_size), _lib.ptr(s_starts),
                                                _lib.ptr(t_starts), _lib.ptr(s_idx), _lib.ptr(t_idx), ctypes.byref(o_ns),
                                                ctypes.byref(o_nt), ctypes.byref(o_ng), groups, _lib.cur_stream()))
    assert o_ns.value == n_s and o_nt.value == n_t
    return s_starts, t_starts, s_idx, t_idx, [(groups[3 * g], groups[3 * g + 1], groups[3 * g + 2])
                                               for g in range(o_ng.value)]


def _dummy_losses(model, dev, dino_loss, scorer, local_clip_size, global_clip_size):
    """Malformed video (frames that are not 224 x 224 after the crop) -> the reference's dummy views: `global_clip_size` identical losses of the all-zero
    (T=local, T=60) pair (dino_loss_loader.py:36-38,103-107; dino_similarity.py:78)."""
    zs = torch.zeros(local_clip_size, 3, 224, 224, device=dev)
    zt = torch.zeros(60, 3, 224, 224, device=dev)
    fs = _forward_windows(model, zs, [0], local_clip_size, 1)
    ft = _forward_windows(model, zt, [0], 60, 1)
    idx = torch.zeros(global_clip_size, dtype=torch.int32, device=dev)
    if scorer is not None:
        return scorer.pair_losses(fs, ft, idx, idx)
    return dino_loss.pair_losses(fs, ft, idx, idx)


def score_videos_device(model, videos, dino_loss, local_clip_size=3, global_clip_size=30, scorer=None,
                        max_rows=1 << 19, small_batch_rows=150000):
    """Scores a BATCH of videos on the GPU; returns one CUDA tensor [F_v] of per-frame losses per video (no host sync).

    model: dvst_b200 VisionTransformer; videos: list of fp32 [F_v,3,H,W] CUDA tensors (normalised, cropped - what the
    reference's loader produces before view construction, dino_loss_loader.py:80-92); dino_loss: DINOLoss (mode R) -
    ignored when a HeadScorer is given (mode H).

    Every video is independent in the reference (dino_similarity.py:66-93 walks them one by one).  Here the student
    windows of all videos (all T = local) go through the backbone as one batch, the teacher windows as one batch per
    distinct G = min(global, F_v), and all frames' pair losses come from one launch.  A clip's features do not depend
    on what else is in its batch (rows of every GEMM / attention tile are independent), so the losses are bit-identical
    to scoring the videos one at a time."""
    videos = list(videos)
    if not videos:
        return []
    dev = videos[0].device
    for v in videos:
        if not v.is_cuda:
            raise _lib.DvstError("dvst_b200 runs on CUDA (sm_100a) only - there is no CPU fallback")
        if v.device != dev:
            raise ValueError("all videos of a batch must live on the same device")
    out = [None] * len(videos)
    with torch.cuda.device(dev):
        ok_shape = [v.dim() == 4 and v.shape[0] >= 1 and v.shape[2] == 224 and v.shape[3] == 224 for v in videos]
        good = [i for i, v in enumerate(videos) if ok_shape[i] and v.shape[0] >= local_clip_size]
        short = [i for i, v in enumerate(videos) if ok_shape[i] and v.shape[0] < local_clip_size]
        if len(good) + len(short) < len(videos):
            dummy = _dummy_losses(model, dev, dino_loss, scorer, local_clip_size, global_clip_size)
            for i in range(len(videos)):
                if not ok_shape[i]:
                    out[i] = dummy
        for i in short:
            # fewer frames than the local window: the reference's broadcast views (short_video_clips), F equal losses
            v = videos[i].detach().to(torch.float32)
            F = v.shape[0]
            si, ti = short_video_clips(F, local_clip_size)
            fs = _forward_windows(model, v[si].contiguous(), [0], F, 1)
            ft = _forward_windows(model, v[ti].contiguous(), [0], F, 1)
            idx = torch.zeros(F, dtype=torch.int32, device=dev)
            out[i] = scorer.pair_losses(fs, ft, idx, idx) if scorer is not None else dino_loss.pair_losses(fs, ft, idx, idx)
        if not good:
            return out
        clips = [videos[i].detach().to(torch.float32).contiguous() for i in good]
        frames = clips[0] if len(clips) == 1 else torch.cat(clips, 0)
        H, W = frames.shape[2], frames.shape[3]
        N = (H // 16) * (W // 16)
        lengths = [c.shape[0] for c in clips]
        # window starts in units of the concatenated frame axis (windows never cross a video boundary: every start is
        # clamped to [0, F_v - T] inside its own video); built on the device, no host round trip
        s_starts, t_starts, s_idx_t, t_idx_t, groups = plan_windows(lengths, local_clip_size, global_clip_size, dev)
        ft = torch.empty(t_starts.numel(), model.embed_dim, dtype=torch.float32, device=dev)
        fs = torch.empty(s_starts.numel(), model.embed_dim, dtype=torch.float32, device=dev)
        rows = N * (s_starts.numel() * local_clip_size + sum(G * cnt for G, cnt, _ in groups))
        # Small batches (a single short video: ~28 three-frame student clips + one 30-frame teacher clip) leave most SMs
        # idle and every kernel latency-bound, and the student and teacher passes are independent: run the teacher pass on
        # a second stream with its own workspace so the two fill the machine together.  Same kernels on the same data:
        # identical losses.  Large batches keep the single stream (nothing to gain, and one workspace).
        overlap = rows <= small_batch_rows
        cur = torch.cuda.current_stream()
        if overlap:
            if getattr(model, "_side_stream", None) is None or model._side_stream.device != dev:
                model._side_stream = torch.cuda.Stream(device=dev)
            side = model._side_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for G, cnt, row0 in groups:
                    _forward_windows(model, frames, t_starts[row0:row0 + cnt], G, max(1, max_rows // (N * G)),
                                     out=ft[row0:row0 + cnt], ws="_workspace_side")
        _forward_windows(model, frames, s_starts, local_clip_size, max(1, max_rows // (N * local_clip_size)), out=fs)
        if overlap:
            cur.wait_stream(side)
        else:
            for G, cnt, row0 in groups:
                _forward_windows(model, frames, t_starts[row0:row0 + cnt], G, max(1, max_rows // (N * G)),
                                 out=ft[row0:row0 + cnt])
        if scorer is not None:
            loss = scorer.pair_losses(fs, ft, s_idx_t, t_idx_t)
        else:
            loss = dino_loss.pair_losses(fs, ft, s_idx_t, t_idx_t)
        for i, piece in zip(good, torch.split(loss, lengths)):
            out[i] = piece
    return out


def score_video_device(model, frames, dino_loss, local_clip_size=3, global_clip_size=30, scorer=None,
                       max_rows=1 << 19, **kw):
    """Scores one video on the GPU; returns the per-frame losses as a CUDA tensor [F] (no host sync).  See
    score_videos_device (this is the batch of one)."""
    if not frames.is_cuda:
        raise _lib.DvstError("dvst_b200 runs on CUDA (sm_100a) only - there is no CPU fallback")
    return score_videos_device(model, [frames], dino_loss, local_clip_size, global_clip_size, scorer, max_rows, **kw)[0]


def score_video(model, frames, dino_loss=None, local_clip_size=3, global_clip_size=30, scorer=None, **kw):
    """Python list of per-frame losses (what the reference appends via `.item()`, dino_similarity.py:87-88)."""
    if dino_loss is None and scorer is None:
        dino_loss = DINOLoss(out_dim=768, teacher_temp=0.02, student_temp=0.3).to(frames.device)
    return score_video_device(model, frames, dino_loss, local_clip_size, global_clip_size, scorer, **kw).tolist()


def dino_similarity(cfg, local_clip_size, global_clip_size, sampling_rate, file_path, dataset=None, model=None,
                    scorer=None, videos_per_batch=8):
    """Driver with the reference's signature and call form (dino_similarity.py:16-93):
    `dino_similarity(args, 3, 30, 4, "loss.json")` with `args` the reference's argparse namespace (`cfg_file`, `opts`,
    `pretrained_weights`, ...) works unchanged.

    Like the reference, `config = load_config(cfg)` resolves the YAML + `--opts` (config.load_config; an already
    resolved config object is used as it is), the model is get_vit_base_patch16_224 + the "backbone."-stripped
    checkpoint `cfg.pretrained_weights`, and the videos are those listed in `<DATA.PATH_TO_DATA_DIR>/test.csv`, decoded
    with torchvision.io.read_video (config.DecodedVideoDataset = the front half of DinoLossLoader).  What differs is where
    the work happens: decoded uint8 frames are uploaded once, sub-sampled / normalised / cropped on the device
    (preprocess_frames), windows are built by index, `videos_per_batch` videos are scored per backbone pass
    (score_videos_device: same losses as one at a time) and each video's F-float loss vector comes back in one D2H copy;
    the JSON is merged once per batch (export_losses: same file content as the reference's per-video export_loss).

    Optional extensions: `dataset` = any iterable of (frames, file_name) with frames either decoded uint8 [n,H,W,3] (not
    yet sub-sampled) or fp32 [F,3,224,224] (already sub-sampled, normalised, cropped); `model` = a prebuilt dvst_b200
    VisionTransformer; `scorer` = a HeadScorer for mode H."""
    from .config import DecodedVideoDataset, load_config
    from .models import get_vit_base_patch16_224
    config = None
    if dataset is None or model is None:
        config = load_config(cfg)
    if dataset is None:
        dataset = DecodedVideoDataset(config, "test")
    if model is None:
        # Mode R (the reference's own loss: 768-d features, tau_t = 0.02) amplifies feature error ~50x, so the model the
        # driver builds for it runs in precision="high" (per-frame loss within 1e-3 of the reference); mode H (a
        # HeadScorer) meets the same bound with bf16 operands.  DVST_PRECISION overrides; a model passed in keeps its own.
        prec = os.environ.get("DVST_PRECISION") or ("high" if scorer is None else "bf16")
        model = get_vit_base_patch16_224(cfg=config, no_head=False, precision=prec)
        if getattr(cfg, "pretrained_weights", None):
            ckpt = torch.load(cfg.pretrained_weights, map_location="cpu")
            renamed = {x[len("backbone."):]: y for x, y in ckpt.items() if x.startswith("backbone.")}
            model.load_state_dict(renamed, strict=False)
        model = model.eval().cuda()
    dev = next(model.parameters()).device
    dino_loss = DINOLoss(out_dim=768, teacher_temp=0.02, student_temp=0.3).to(dev)
    results = {}

    def flush(batch):
        with torch.no_grad():
            losses = score_videos_device(model, [f for f, _ in batch], dino_loss, local_clip_size, global_clip_size,
                                         scorer=scorer)
        out = {}
        for (_, file_name), loss in zip(batch, losses):
            out[file_name] = loss.tolist()
            results[os.path.splitext(os.path.basename(file_name))[0]] = out[file_name]
        export_losses(out, file_path)

    mean = getattr(getattr(config, "DATA", None), "MEAN", (0.45, 0.45, 0.45))
    std = getattr(getattr(config, "DATA", None), "STD", (0.225, 0.225, 0.225))
    batch = []
    for item in dataset:
        frames, file_name = item[0], item[1]
        if isinstance(file_name, (list, tuple)):   # a DataLoader collates the path into a 1-list (dino_similarity.py:93)
            file_name = file_name[0]
        frames = frames.to(dev, non_blocking=True)
        if frames.dtype == torch.uint8:
            if frames.dim() == 5:                  # DataLoader batch dimension of 1
                frames = frames[0]
            frames = preprocess_frames(frames, sampling_rate, mean, std)
        batch.append((frames, file_name))
        if len(batch) >= max(1, int(videos_per_batch)):
            flush(batch)
            batch = []
    if batch:
        flush(batch)
    return results
