"""Multi-GPU plumbing for the scoring path: one process per GPU (torchrun), videos sharded across ranks, no data-path
collective; NCCL (gloo on CPU in the tests) only for the gather of per-video score vectors and, in the k-NN feature
extraction, the all-gather of CLS features (eval_knn.py:111-132).  SURVEY.md section 8e.
"""
import torch
import torch.distributed as dist


def video_cost(F, local_clip_size=3, global_clip_size=30):
    """Relative cost of scoring a video with F sampled frames: distinct student windows x T=local + distinct teacher
    windows x T=G (FLOPs are ~linear in T)."""
    if F < local_clip_size:
        return 2 * max(F, 1)   # one student + one teacher clip of F frames (dino_similarity.short_video_clips)
    G = min(global_clip_size, F)
    return (F - local_clip_size + 1) * local_clip_size + (F - G + 1) * G


def shard_videos(lengths, world_size, local_clip_size=3, global_clip_size=30):
    """Greedy longest-processing-time partition of video indices over ranks (deterministic on every rank).
    Returns a list (per rank) of lists of video indices."""
    order = sorted(range(len(lengths)), key=lambda i: (-video_cost(lengths[i], local_clip_size, global_clip_size), i))
    loads = [0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += video_cost(lengths[i], local_clip_size, global_clip_size)
    for p in parts:
        p.sort()
    return parts


def gather_scores(local_scores, lengths, parts, device):
    """All ranks -> rank 0: `local_scores` maps video index -> 1-D float tensor of per-frame losses (this rank's
    shard).  One padded [n_local, F_max] all_gather instead of a per-frame `.item()` (dino_similarity.py:87-88).
    Returns {video index: list of floats} on rank 0, None elsewhere."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return {i: local_scores[i].tolist() for i in parts[0]}

    def width(F):
        return max(F, 30)  # dummy (wrong-crop) videos produce 30 values
    n_max = max(len(p) for p in parts)
    f_max = max(width(f) for f in lengths) if lengths else 1
    buf = torch.zeros(n_max, f_max, dtype=torch.float32, device=device)
    cnt = torch.zeros(n_max, dtype=torch.int64, device=device)
    for k, i in enumerate(parts[rank]):
        v = local_scores[i].to(device=device, dtype=torch.float32)
        buf[k, : v.numel()] = v
        cnt[k] = v.numel()
    bufs = [torch.empty_like(buf) for _ in range(world)]
    cnts = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(bufs, buf)
    dist.all_gather(cnts, cnt)
    if rank != 0:
        return None
    out = {}
    for r in range(world):
        b, c = bufs[r].cpu(), cnts[r].cpu()
        for k, i in enumerate(parts[r]):
            out[i] = b[k, : int(c[k])].tolist()
    return out


@torch.no_grad()
def extract_features(model, batches, n_total, device):
    """eval_knn.extract_features (eval_knn.py:94-135): frozen-backbone CLS features of a sharded dataset gathered
    into a [n_total, 768] matrix on rank 0.  `batches` yields (samples [B,3,T,H,W], index [B] int64)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    features = None
    for samples, index in batches:
        feats = model(samples.to(device, non_blocking=True)).clone()
        index = index.to(device)
        if rank == 0 and features is None:
            features = torch.zeros(n_total, feats.shape[-1], device=device)
        if world > 1:
            idx_all = [torch.empty_like(index) for _ in range(world)]
            f_all = [torch.empty_like(feats) for _ in range(world)]
            dist.all_gather(idx_all, index)
            dist.all_gather(f_all, feats)
            index, feats = torch.cat(idx_all), torch.cat(f_all)
        if rank == 0:
            features.index_copy_(0, index, feats)
    return features
