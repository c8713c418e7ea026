"""The whole summarisation path on N GPUs, the way a driver would run it (torchrun, one process per GPU, NCCL):

  synthetic decoded videos (uint8 [n, 224, 224, 3], host) -> sharded over ranks by cost (parallel.shard_videos)
  -> upload as uint8 + fused sub-sample / normalise / transpose (preprocess_frames) -> batched whole-video scoring
  (score_videos_device) -> per-video loss vectors gathered on rank 0 (parallel.gather_scores, one padded all-gather)
  -> the reference's JSON (export_losses) -> adaptive frame selection for every video in one launch
  (select_frames_adaptive_batch).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/pipeline_demo.py \
        --videos 12 --depth 2 --out /tmp/loss.json --check

--check recomputes every video alone on rank 0 and compares (bit-identical losses, identical index lists), then prints
`pipeline ok`.  Synthetic weights / videos (no datasets or checkpoints offline)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dvst_b200  # noqa: E402,F401
from dvst_b200.dino_similarity import DINOLoss, export_losses, preprocess_frames, score_videos_device  # noqa: E402
from dvst_b200.frame_selection import select_frames_adaptive, select_frames_adaptive_batch  # noqa: E402
from dvst_b200.models.timesformer import VisionTransformer  # noqa: E402
from dvst_b200.parallel import gather_scores, shard_videos  # noqa: E402
from dvst_b200.testing import synth_backbone_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=12)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--out", default="/tmp/dvst_loss.json")
    ap.add_argument("--frames", type=int, default=8, help="frames to select per video")
    ap.add_argument("--videos-per-pass", type=int, default=4)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from functools import partial
    m = VisionTransformer(img_size=224, num_classes=0, depth=a.depth, qkv_bias=True,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
    m.load_state_dict(synth_backbone_state_dict(0, depth=a.depth, num_frames=8), strict=True)
    m = m.eval().to(dev)
    dl = DINOLoss(768, 0.02, 0.3).to(dev)

    # "decoded" videos: every rank derives the same list (lengths follow the MSVD-like spread of the shipped loss file)
    g = torch.Generator().manual_seed(123)
    n_raw = [int(v) for v in torch.randint(44, 280, (a.videos,), generator=g)]
    n_raw[0] = 7            # -> 2 sampled frames: shorter than the local window (the reference's broadcast views)
    lengths = [-(-n // 4) for n in n_raw]
    names = [f"/data/video_{i:04d}.avi" for i in range(a.videos)]

    def raw_video(i):
        gi = torch.Generator().manual_seed(1000 + i)
        return torch.randint(0, 256, (n_raw[i], 224, 224, 3), generator=gi, dtype=torch.uint8)

    parts = shard_videos(lengths, world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scores = {}
    mine = parts[rank]
    with torch.no_grad():
        for k in range(0, len(mine), a.videos_per_pass):
            idx = mine[k:k + a.videos_per_pass]
            frames = [preprocess_frames(raw_video(i).pin_memory().to(dev, non_blocking=True), 4) for i in idx]
            for i, s in zip(idx, score_videos_device(m, frames, dl)):
                scores[i] = s
    all_scores = gather_scores(scores, lengths, parts, dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ok = True
    if rank == 0:
        if os.path.exists(a.out):
            os.remove(a.out)
        export_losses({names[i]: all_scores[i] for i in range(a.videos)}, a.out)
        written = json.load(open(a.out))
        keys = [os.path.splitext(os.path.basename(n))[0] for n in names]
        sel = select_frames_adaptive_batch([written[k] for k in keys], n_raw, 4, a.frames, device=dev).cpu().tolist()
        print(f"scored {a.videos} videos ({sum(lengths)} sampled frames) on {world} GPU(s) in {dt:.2f} s "
              f"({a.videos / dt:.1f} videos/s incl. synthetic decode + upload); JSON -> {a.out}")
        if a.check:
            with torch.no_grad():
                for i in range(a.videos):
                    alone = score_videos_device(m, [preprocess_frames(raw_video(i).to(dev), 4)], dl)[0].tolist()
                    ok &= alone == written[keys[i]]
                    ok &= len(alone) == lengths[i]
                    ok &= select_frames_adaptive(alone, n_raw[i], 4, a.frames) == sel[i]
            print("pipeline ok" if ok else "pipeline MISMATCH")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
