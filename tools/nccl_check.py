"""World-size-N NCCL check of the two collectives of the path that the reference runs through torch.distributed, with the
REAL model on real GPUs (the gloo / CPU tests in tests/test_distributed_cpu.py use a fake model):

  * eval_knn.extract_features (eval_knn.py:94-135): sharded batches -> backbone -> all_gather of indices and CLS
    features -> rank 0 holds [n_total, 768]; must equal the single-GPU forward of the whole set bit for bit
    (a clip's features do not depend on its batch);
  * train_ssl.DINOLoss.update_center (train_ssl.py:673-691): per-rank teacher batch -> column sum -> all_reduce ->
    / (B * world) -> EMA; must equal the single-process update on the concatenated teacher batch;
  * parallel.gather_scores of per-video loss vectors scored on each rank's shard.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_check.py

Prints `nccl ok` on rank 0 and exits 0 when everything matches."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dvst_b200  # noqa: E402,F401
from dvst_b200.dino_similarity import DINOLoss, score_videos_device  # noqa: E402
from dvst_b200.models.timesformer import VisionTransformer  # noqa: E402
from dvst_b200.parallel import extract_features, gather_scores, shard_videos  # noqa: E402
from dvst_b200.testing import synth_backbone_state_dict, synth_clips, synth_video  # noqa: E402
from dvst_b200.train_loss import DINOLoss as TrainDINOLoss  # noqa: E402


def main():
    from functools import partial
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    m = VisionTransformer(img_size=224, num_classes=0, depth=2, qkv_bias=True,
                          norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
    m.load_state_dict(synth_backbone_state_dict(1, depth=2, num_frames=8), strict=True)
    m = m.eval().to(dev)
    ok = True

    # ---- extract_features: 12 clips of 8 frames, strided over the ranks in batches of 3 (DistributedSampler order) ----
    n_total = 12
    data = synth_clips(31, n_total, 8)
    idx_all = torch.arange(n_total)
    mine = idx_all[rank::world]
    batches = [(data[mine[k:k + 3]], mine[k:k + 3]) for k in range(0, len(mine), 3)]
    feats = extract_features(m, batches, n_total, dev)
    if rank == 0:
        with torch.no_grad():
            ref = m(data.to(dev))
        ok &= bool(torch.equal(feats, ref))
        print("extract_features equal to single-GPU forward:", bool(torch.equal(feats, ref)))

    # ---- training DINOLoss.update_center: all_reduce of the batch centre ----
    out_dim, B = 4096, 3
    g = torch.Generator().manual_seed(5)
    teacher_all = torch.randn(world, 2 * B, out_dim, generator=g)
    student_all = torch.randn(world, 4 * B, out_dim, generator=g)
    c0 = torch.randn(1, out_dim, generator=g) * 0.1
    L = TrainDINOLoss(out_dim, 4, 0.04, 0.07, 3, 10).to(dev)
    L.center.copy_(c0)
    loss = L(student_all[rank].to(dev), teacher_all[rank].to(dev), epoch=1)
    expect = 0.9 * c0 + 0.1 * teacher_all.reshape(-1, out_dim).sum(0, keepdim=True) / (2 * B * world)
    err = (L.center.cpu() - expect).abs().max().item()
    centers = [torch.empty_like(L.center) for _ in range(world)]
    dist.all_gather(centers, L.center)
    same = all(torch.equal(c, centers[0]) for c in centers)
    ok &= err < 1e-6 and same and bool(torch.isfinite(loss))
    if rank == 0:
        print(f"update_center: max err vs concatenated-batch update {err:.2e}, identical on all ranks: {same}")

    # ---- gather_scores: per-video loss vectors of each rank's shard -> rank 0 ----
    lengths = [5, 9, 2, 12, 4, 7]
    parts = shard_videos(lengths, world, 3, 6)
    dl = DINOLoss(768, 0.02, 0.3).to(dev)
    vids = {i: synth_video(40 + i, lengths[i], smooth=0.5).to(dev) for i in parts[rank]}
    with torch.no_grad():
        sc = dict(zip(parts[rank], score_videos_device(m, [vids[i] for i in parts[rank]], dl, 3, 6)))
    allsc = gather_scores(sc, lengths, parts, dev)
    if rank == 0:
        with torch.no_grad():
            for i, F in enumerate(lengths):
                alone = score_videos_device(m, [synth_video(40 + i, F, smooth=0.5).to(dev)], dl, 3, 6)[0].tolist()
                ok &= alone == allsc[i]
        print("gather_scores equal to single-GPU scoring:", ok)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("nccl ok" if int(flag) == 1 else "nccl MISMATCH")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
