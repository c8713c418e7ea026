"""N > 1 host path on CPU: world_size-2 gloo.  Videos are sharded with no data-path collective; only the per-video
score vectors (and k-NN features) are gathered."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dvst_b200  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import dvst_b200  # noqa: F401
    from dvst_b200.parallel import extract_features, gather_scores, shard_videos
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = [30, 68, 11, 5, 2, 40, 33]
    parts = shard_videos(lens, world)
    local = {}
    for i in parts[rank]:
        n = lens[i] if lens[i] >= 3 else 30
        local[i] = torch.arange(n, dtype=torch.float32) + 1000 * i
    got = gather_scores(local, lens, parts, torch.device("cpu"))
    ok = True
    if rank == 0:
        for i, f in enumerate(lens):
            n = f if f >= 3 else 30
            ok &= got[i] == [float(v + 1000 * i) for v in range(n)]
    else:
        ok &= got is None

    class Fake(torch.nn.Module):
        def forward(self, x):
            return x.flatten(1)[:, :8].float()
    n_total = 12
    idx_all = torch.arange(n_total)
    mine = idx_all[rank::world]
    data = torch.arange(n_total * 16, dtype=torch.float32).reshape(n_total, 1, 1, 4, 4)
    batches = [(data[mine[k:k + 3]], mine[k:k + 3]) for k in range(0, len(mine), 3)]
    feats = extract_features(Fake(), batches, n_total, torch.device("cpu"))
    if rank == 0:
        ok &= torch.equal(feats, data.flatten(1)[:, :8])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_scores_and_features_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == {0: True, 1: True}
