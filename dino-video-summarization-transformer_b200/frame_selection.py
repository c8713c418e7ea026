"""Adaptive frame selection from the scoring JSON - the consumer of the hot path's output
(datasets_custom/frame_selection_loader.py:115-190, duplicated in testing/extract_video_adaptive.py:43-98).

Pure host logic (numpy, fp64) exactly like the reference: inverse-CDF sampling of N quantiles of the normalised loss
curve, indices scaled by the pre-sampling rate, collisions resolved by walking forward `rate` frames at a time.
"""
import json

import numpy as np


def select_frames_adaptive(loss_list, n_unsampled, pre_sampling_rate, num_frames, n_sampled=None):
    """Returns `num_frames` frame indices (in unsampled-video units).

    loss_list: per-sampled-frame losses; n_unsampled: frames in the original video (frames_unsampled.size(0));
    n_sampled: frames after pre-sampling (defaults to ceil(n_unsampled / rate)); the loss list is truncated to it
    (frame_selection_loader.py:125-126)."""
    if n_sampled is None:
        n_sampled = -(-n_unsampled // pre_sampling_rate)
    loss = np.asarray(loss_list)
    if len(loss) > n_sampled:
        loss = loss[:n_sampled]
    pdf = loss / np.sum(loss)   # the min-max line (:129) is dead code in the reference: overwritten by this one
    cdf = np.cumsum(pdf)
    indices = []
    for i in range(num_frames):
        idx = int(np.abs(cdf - i / num_frames).argmin())
        scaled = idx * pre_sampling_rate
        if scaled not in indices:
            indices.append(scaled)
            continue
        temp = scaled + pre_sampling_rate
        while True:
            if temp not in indices and temp < n_unsampled:
                indices.append(temp)
                break
            if temp >= n_unsampled:
                indices.append(temp - pre_sampling_rate)
                break
            temp += pre_sampling_rate
    return indices


def select_frames_adaptive_batch(loss_lists, n_unsampled, pre_sampling_rate, num_frames, device="cuda"):
    """select_frames_adaptive for a whole batch of videos in one kernel launch (dvst_select_frames, csrc/selection.cu):
    one device thread per video in the reference's fp64 arithmetic, so every index list equals the host function's.

    loss_lists: one loss sequence per video (as stored by export_loss); n_unsampled: frames of each undecimated video.
    Returns an int32 CUDA tensor [V, num_frames] of frame indices in unsampled-video units."""
    import torch

    from . import _lib
    V = len(loss_lists)
    n_samp = [-(-int(nu) // pre_sampling_rate) for nu in n_unsampled]
    lens = [min(len(l), ns) for l, ns in zip(loss_lists, n_samp)]
    fmax = max(lens) if lens else 1
    host = torch.zeros(V, fmax, dtype=torch.float64)
    for v, l in enumerate(loss_lists):
        host[v, :lens[v]] = torch.as_tensor(np.asarray(l, dtype=np.float64)[:lens[v]])
    dev = torch.device(device)
    loss = host.to(dev)
    n_loss = torch.tensor(lens, dtype=torch.int32, device=dev)
    n_un = torch.tensor([int(x) for x in n_unsampled], dtype=torch.int32, device=dev)
    out = torch.empty(V, num_frames, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dvst_select_frames(_lib.ptr(loss), fmax, _lib.ptr(n_loss), _lib.ptr(n_un), V,
                                                 pre_sampling_rate, num_frames, _lib.ptr(out), _lib.cur_stream()))
    return out


def select_frames_uniform(n_sampled, num_frames):
    """Uniform baseline (frame_selection_loader.py:175-185): indices in SAMPLED-frame units, like the reference."""
    interval = int(n_sampled / num_frames)
    return [i * interval for i in range(num_frames)]


def select_from_json(loss_file, video_name, n_unsampled, pre_sampling_rate=4, num_frames=8):
    with open(loss_file, "r") as f:
        loss_dict = json.load(f)
    return select_frames_adaptive(loss_dict[video_name], n_unsampled, pre_sampling_rate, num_frames)
