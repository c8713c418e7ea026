#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native scoring path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Metric (BASELINE.json): videos/sec scored.  Workload = BASELINE.json configs[1]: TimeSformer ViT-B/16, 16 frames,
224^2, Kinetics-400-shaped synthetic clips, batch 64 per GPU.  One step = 64 clips (1 clip = 1 video) through the
backbone forward, the DINO projection head and the fused centred / sharpened cross-entropy on the 32 (student, teacher)
pairs formed by (clip 2i, clip 2i+1); weak scaling (every rank scores its own 64 videos, no data-path collective; the
32 per-rank losses are all-gathered).  `value` is timed with inputs resident in HBM; `e2e` goes through the public
module API from pinned HOST buffers: the step's 1 024 decoded frames are uploaded as uint8 (what a decoder produces),
normalised, windowed into clips by index and cut into patch rows in one pass on the device (dvst_forward_windows_u8:
the loader front half inside the patch extraction, bit-identical to dvst_preprocess_u8 + dvst_forward_windows), and the
losses are read back - H2D and D2H inside the timed region.

Beside the headline the line carries
  * `parity`      - the GPU step's features / loss for two cfg2 clips against the CPU arm's (the run FAILS above 2e-2 / 1e-3),
  * `roofline`    - the tensor-bound GEMM launches, from CUDA events inside the timed region; executed FLOPs only,
  * `aux.cfg3_msvd_*` - BASELINE configs[2]: whole-video 4_3_30 scoring of MSVD-shaped videos, end to end (uint8 host
                    frames -> H2D -> preprocess -> windows -> backbone -> loss -> gather -> JSON -> frame selection), sharded
                    over the ranks with parallel.shard_videos, in mode H (bf16) and in the reference's own mode R
                    (precision="high"),
  * `cpu_baseline` / `--impl reference` - the UNMODIFIED reference modules (staged under oracle/_ref by build()) on the
                    host cores; the oracle port only when that copy is absent.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "videos_per_sec_scored"
UNIT = "videos/s"
B_PER_GPU, T, HW = 64, 16, 224
OUT_DIM = 65536
TT, TS = 0.04, 0.1
PARITY_FEAT_TOL, PARITY_LOSS_TOL = 2e-2, 1e-3


def clip_gflop(t, n=196):
    """Algorithmic GFLOP of one clip's backbone forward in the reference formulation."""
    s = t * n
    g = 2 * s * 768 * 2304 + 2 * (2 * s * 768 * 768) + 2 * t * (n + 1) * 768 * 2304 + 2 * t * (n + 1) * 768 * 768 + \
        2 * 2 * (s + 1) * 768 * 3072
    a = 4 * n * t * t * 768 + 4 * t * (n + 1) ** 2 * 768
    return (2 * t * n * 768 * 768 + 12 * (g + a)) / 1e9


GF_PER_CLIP_ALGO = clip_gflop(T)   # 784.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary config-3 (4_3_30 video) numbers")
    ap.add_argument("--msvd-videos", type=int, default=48, help="MSVD-shaped videos per GPU in the cfg3 leg (mode H)")
    return ap.parse_args()


def workload_config(n_gpus):
    return {"workload": "cfg2: TimeSformer ViT-B/16 divided space-time, 64 clips x 16 frames x 224^2 per GPU per step "
                        "(Kinetics-400-shaped, 1 clip = 1 video): backbone forward + DINOHead(768->65536) + fused "
                        "centred/sharpened CE on 32 (student, teacher) pairs",
            "clips_per_gpu": B_PER_GPU, "frames": T, "crop": HW, "pairs_per_gpu": B_PER_GPU // 2, "out_dim": OUT_DIM,
            "teacher_temp": TT, "student_temp": TS, "weights": "random init, seed 0 (all tensors randomised)",
            "parallelism": f"dp{n_gpus} (videos sharded, replicated frozen weights)",
            "l2_policy": "inputs exceed L2 (616 MB of clips per step vs 126 MB L2); activations 3 GB per step",
            "precision": "bf16 operands, fp32 accumulate / residual / LayerNorm / softmax",
            "e2e_input": "uint8 decoded frames [1024,224,224,3] from pinned host memory (154 MB per step), normalised on "
                         "the device inside the patch extraction (dvst_forward_windows_u8)"}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference modules (oracle/_ref staged by build(), or /root/reference) on
# the host cores; the oracle port only when neither is present
# ---------------------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self):
        import torch
        from functools import partial
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import dvst_b200  # noqa: F401
        from dvst_b200.testing import synth_backbone_state_dict, synth_head_state_dict
        self.torch = torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = synth_backbone_state_dict(0, depth=12, num_frames=8)
        self.hsd = synth_head_state_dict(2, out_dim=OUT_DIM)
        import ref_shim
        self.kind = "port"
        if ref_shim.available():
            try:
                R = ref_shim.load()
                m = R["VisionTransformer"](img_size=224, num_classes=0, patch_size=16, embed_dim=768, depth=12,
                                           num_heads=12, mlp_ratio=4, qkv_bias=True,
                                           norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8,
                                           attention_type="divided_space_time")
                m.load_state_dict(self.sd, strict=True)
                head = R["DINOHead"](768, OUT_DIM)
                head.load_state_dict(self.hsd, strict=True)
                self.model, self.head = m.eval(), head.eval()
                self.loss = R["DINOLoss"](out_dim=OUT_DIM, teacher_temp=TT, student_temp=TS)
                self.kind = "reference"
                self.where = ref_shim.REF_ROOT
            except Exception as e:  # pragma: no cover - falls back to the port, and says so
                print(f"[bench] reference modules unusable ({e!r}); timing the oracle port instead", file=sys.stderr)
        if self.kind == "port":
            import dvst_oracle as O
            self.O = O
            self.where = "oracle/dvst_oracle.py"

    def step(self, clips):
        """The cfg2 step on `clips` [B,3,T,224,224] (B even): (features [B,768], per-pair losses [B/2])."""
        torch = self.torch
        with torch.no_grad():
            if self.kind == "reference":
                f = self.model(clips)
                logits = self.head(f)
                loss = torch.stack([self.loss(logits[2 * k], logits[2 * k + 1]) for k in range(clips.shape[0] // 2)])
            else:
                f = self.O.forward_features(self.sd, clips)
                logits = self.O.dino_head(self.hsd, f)
                loss = self.O.dino_pair_loss(logits[0::2], logits[1::2], torch.zeros(1, OUT_DIM), TT, TS)
        return f, loss

    def describe(self, clips, steps):
        what = ("the UNMODIFIED reference modules (models/timesformer.py VisionTransformer, vision_transformer.py DINOHead, "
                "dino_similarity.py DINOLoss) imported from " + self.where) if self.kind == "reference" else \
            "oracle/dvst_oracle.py = CPU restatement of the reference's PyTorch path (no staged reference copy found)"
        return f"{clips} clips (1 pair) of the cfg2 step per timed step, {steps} timed steps; {what}; torch CPU fp32"


def run_cpu_arm(n_steps, n_warm, clips_per_step):
    arm = CpuArm()
    from dvst_b200.testing import synth_clips
    clips = synth_clips(1, clips_per_step, T)
    f = loss = None
    for _ in range(n_warm):
        f, loss = arm.step(clips)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        f, loss = arm.step(clips)
    dt = time.perf_counter() - t0
    return arm, clips_per_step * n_steps / dt, dt / n_steps, f, loss


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = 2
    # K timed and W warm-up steps as asked (one step = the 2-clip sample, about 1 s on 16 host cores); the caps only keep an
    # absurd request inside a few minutes
    steps, warm = max(1, min(args.steps, 60)), max(1, min(args.warmup, 10))
    arm, vps, s_per_step, _, _ = run_cpu_arm(steps, warm, clips)
    line = {"impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": vps, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                             "sample": arm.describe(clips, steps)},
            "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    # started BEFORE the warm-up (nvidia-smi needs ~0.3 s to deliver its first sample); samples are then filtered to the
    # timed window by their timestamps
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t_start=None, t_end=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        if t_start is not None and t_end is not None:
            def ts(r):
                try:
                    return datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except Exception:
                    return None
            inside = [r for r in rows if ts(r) is not None and t_start <= ts(r) <= t_end + 0.05]
            if len(inside) >= 2:
                rows = inside
        sm, reasons, mx, pw = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                r = [c.strip() for c in r]
                sm.append(float(r[1]))
                mx = float(r[2])
                pw.append(float(r[3]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            # "under load" = samples in the upper half of the observed power range
            thr = (max(pw) + min(pw)) / 2 if pw else 0
            load = [s for s, p_ in zip(sm, pw) if p_ >= thr] or sm
            load.sort()
            out = {"sm_mhz": load[len(load) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(pw) if pw else None}
        return out


# ---------------------------------------------------------------------------------------------------------------------
# cfg3: whole-video 4_3_30 scoring of MSVD-shaped videos, end to end and sharded over the ranks
# ---------------------------------------------------------------------------------------------------------------------
def median_ms(torch, fn, n, sync=None):
    """Median device time of n calls of fn (one CUDA-event pair per call): the auxiliary legs are short, a one-off hiccup of a
    shared box must not end up in their numbers (the headline `value` is timed as one K-step region as the contract says)."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    out = None
    for e0, e1 in ev:
        e0.record()
        out = fn()
        e1.record()
    (sync or torch.cuda.synchronize)()
    t = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return t[len(t) // 2], out


def msvd_lengths(n):
    """n video lengths (sampled frames F) at evenly spaced quantiles of the 1 970 list lengths of the reference's shipped
    loss file (tests/golden/msvd_lengths.json, written by oracle/make_golden.py)."""
    h = json.load(open(os.path.join(ROOT, "tests", "golden", "msvd_lengths.json")))["hist"]
    allv = []
    for k in sorted(h, key=int):
        allv += [int(k)] * int(h[k])
    out = [allv[min(len(allv) - 1, int((i + 0.5) * len(allv) / n))] for i in range(n)]
    # deterministic shuffle so that consecutive videos are not sorted by length
    return [out[(i * 37) % n] for i in range(n)] if n % 37 else out


def video_executed_gflop(F, local=3, glob=30):
    """De-duplicated distinct windows: (F - 2) student clips of 3 frames + max(F - 29, 1) teacher
    clips of G = min(30, F) frames."""
    G = min(glob, F)
    return (F - local + 1) * clip_gflop(local) + (F - G + 1) * clip_gflop(G)


def cfg3_msvd_leg(torch, dist, model, dino_loss, scorer, lengths, world, rank, dev, frames_per_pass=160, select_n=8):
    from dvst_b200.dino_similarity import export_losses, preprocess_frames, score_videos_device
    from dvst_b200.frame_selection import select_frames_adaptive_batch
    from dvst_b200.parallel import gather_scores, shard_videos
    parts = shard_videos(lengths, world)
    mine = parts[rank]
    f_max = max(lengths)
    g = torch.Generator().manual_seed(4321 + rank)
    pool = torch.randint(0, 256, (f_max + 64, HW, HW, 3), generator=g, dtype=torch.uint8).pin_memory()

    def host_video(i):  # decoded + sub-sampled uint8 frames of video i (a window of the random pool), pinned host memory
        off = (i * 13) % (pool.shape[0] - lengths[i] + 1)
        return pool[off:off + lengths[i]]

    def run(ids):
        scores, k = {}, 0
        with torch.no_grad():
            while k < len(ids):
                batch, nfr = [], 0
                while k < len(ids) and (not batch or nfr + lengths[ids[k]] <= frames_per_pass):
                    batch.append(ids[k])
                    nfr += lengths[ids[k]]
                    k += 1
                frames = [preprocess_frames(host_video(i).to(dev, non_blocking=True), 1) for i in batch]
                for i, s in zip(batch, score_videos_device(model, frames, dino_loss, scorer=scorer)):
                    scores[i] = s
        return scores

    # warm-up: workspaces, kernel attributes, every window length class once
    run(sorted(mine, key=lambda i: lengths[i])[:2] + sorted(mine, key=lambda i: -lengths[i])[:1])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    scores = run(mine)
    e1.record()
    torch.cuda.synchronize()
    busy_ms = e0.elapsed_time(e1)
    allsc = gather_scores(scores, lengths, parts, dev)
    sel = None
    if rank == 0:
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "loss.json")
            export_losses({f"/data/video_{i:05d}.avi": allsc[i] for i in range(len(lengths))}, path)
            written = json.load(open(path))
        keys = [f"video_{i:05d}" for i in range(len(lengths))]
        sel = select_frames_adaptive_batch([written[k] for k in keys], [4 * f for f in lengths], 4, select_n,
                                           device=dev).cpu()
        assert sel.shape == (len(lengths), select_n) and all(len(written[k]) == f for k, f in zip(keys, lengths))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([busy_ms, dt * 1e3], dtype=torch.float64, device=dev)
    tmax, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    busy_max, wall_max = float(tmax[0]) / 1e3, float(tmax[1]) / 1e3
    busy_mean = float(tsum[0]) / 1e3 / world
    gf = sum(video_executed_gflop(f) for f in lengths)
    h2d = sum(lengths[i] for i in mine) * HW * HW * 3
    return {"videos": len(lengths), "sampled_frames": sum(lengths), "videos_per_s_device": len(lengths) / busy_max,
            "videos_per_s_e2e": len(lengths) / wall_max, "executed_tflops_per_gpu": gf / 1e3 / busy_max / world,
            "rank_busy_s_max": busy_max, "rank_busy_s_mean": busy_mean, "imbalance_max_over_mean": busy_max / busy_mean,
            "h2d_bytes_rank0": h2d, "d2h": "one padded all-gather of the per-video loss vectors; JSON + selection on rank 0"}


def main_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if world > 1:
        torch.cuda.set_device(local_rank)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    import dvst_b200  # noqa: F401
    from functools import partial
    from dvst_b200 import _lib
    from dvst_b200.dino_similarity import (DINOLoss, HeadScorer, forward_windows_u8, score_video_device,
                                           score_videos_device)
    from dvst_b200.models.timesformer import VisionTransformer
    from dvst_b200.testing import synth_backbone_state_dict, synth_clips, synth_head_state_dict, synth_video
    from dvst_b200.vision_transformer import DINOHead
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    sd = synth_backbone_state_dict(0, depth=12, num_frames=8)
    model = VisionTransformer(img_size=224, num_classes=0, depth=12, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(dev)
    head = DINOHead(768, OUT_DIM)
    head.load_state_dict(synth_head_state_dict(2, out_dim=OUT_DIM), strict=True)
    head = head.to(dev).eval()
    scorer = HeadScorer(head, TT, TS)
    P = B_PER_GPU // 2
    s_idx = torch.arange(0, B_PER_GPU, 2, dtype=torch.int32, device=dev)
    t_idx = torch.arange(1, B_PER_GPU, 2, dtype=torch.int32, device=dev)

    g = torch.Generator().manual_seed(1234 + rank)
    dev_clips = [torch.randn(B_PER_GPU, 3, T, HW, HW, generator=g).to(dev) for _ in range(2)]
    gathered = torch.empty(world * P, dtype=torch.float32, device=dev) if world > 1 else None

    def finish(f):
        loss = scorer.pair_losses(f, f, s_idx, t_idx)
        if world > 1:
            dist.all_gather_into_tensor(gathered, loss)
            return gathered
        return loss

    def step(x):
        with torch.no_grad():
            return finish(model(x))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing (`value`) ----
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])) if rank == 0 else None
    for i in range(args.warmup):
        step(dev_clips[i % 2])
    sync_all()
    _lib.profile_enable(True)
    _lib.profile_collect()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        out = step(dev_clips[i % 2])
    e1.record()
    sync_all()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    prof = _lib.profile_collect()
    _lib.profile_enable(False)
    clocks = sampler.stop(wall0, wall1) if sampler is not None else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * B_PER_GPU * args.steps / (ms_max / 1e3)
    assert torch.isfinite(out).all()

    # ---- parity of the bench configuration itself: two cfg2 clips, GPU step vs the CPU arm (rank 0, N = 1) ----
    gpu_par = None
    if rank == 0:
        xp = synth_clips(1, 2, T).to(dev)
        with torch.no_grad():
            fp = model(xp)
            # batch-composition independence: the same two clips inside a full 64-clip batch give the same bits
            big = dev_clips[0].clone()
            big[10:12] = xp
            fp_in_batch = model(big)[10:12]
            lp = scorer.pair_losses(fp, fp, torch.zeros(1, dtype=torch.int32, device=dev),
                                    torch.ones(1, dtype=torch.int32, device=dev))
        gpu_par = (fp.cpu(), lp.cpu(), bool(torch.equal(fp, fp_in_batch)))
        del big, xp

    # ---- end to end through the public API from pinned host buffers: uint8 frames up, losses down, every step ----
    del dev_clips
    torch.cuda.empty_cache()
    n_frames = B_PER_GPU * T
    gh = torch.Generator().manual_seed(99 + rank)
    host_u8 = [torch.randint(0, 256, (n_frames, HW, HW, 3), generator=gh, dtype=torch.uint8).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty(n_frames, HW, HW, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    host_loss = [torch.empty(world * P if world > 1 else P, dtype=torch.float32).pin_memory() for _ in range(2)]
    starts_dev = torch.arange(0, n_frames, T, dtype=torch.int32, device=dev)

    def step_u8(u8):
        with torch.no_grad():
            # decoded uint8 frames -> normalise + crop + clip windows + patch rows in one pass (dvst_forward_windows_u8)
            return finish(forward_windows_u8(model, u8, starts_dev, T, B_PER_GPU, sampling_rate=1))

    def e2e_run(n_steps):
        results = []
        cur = torch.cuda.current_stream()
        pending = None
        for f_ in free:
            f_.record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[0])
            stage[0].copy_(host_u8[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n_steps):
            b = i % 2
            if i + 1 < n_steps:
                nb = (i + 1) % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[nb])
                    stage[nb].copy_(host_u8[nb], non_blocking=True)
                    ready[nb].record(copy_stream)
            cur.wait_event(ready[b])
            loss = step_u8(stage[b])
            free[b].record(cur)
            host_loss[b].copy_(loss, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cur)
            if pending is not None:
                pending[0].synchronize()
                results.append(float(pending[1].sum()))
            pending = (done, host_loss[b])
        pending[0].synchronize()
        results.append(float(pending[1].sum()))
        return results

    e2e_run(2)
    sync_all()
    t0 = time.perf_counter()
    res = e2e_run(args.steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert len(res) == args.steps and all(r == r for r in res)
    t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B_PER_GPU * args.steps / float(t_e.item())
    h2d = host_u8[0].numel()
    d2h = host_loss[0].numel() * 4
    del host_u8, stage
    torch.cuda.empty_cache()

    # ---- auxiliary: config 3 (4_3_30 adaptive-frame-selection scoring of whole videos) ----
    aux = None
    if not args.no_aux:
        aux = {}
        dl = DINOLoss(768, 0.02, 0.3).to(dev)
        if rank == 0:
            for F in (30, 68):
                frames = synth_video(5, F).to(dev)
                with torch.no_grad():
                    for _ in range(2):       # warm-up with the timed call itself (workspaces, side stream, scratch)
                        score_video_device(model, frames, dl, scorer=scorer)
                    torch.cuda.synchronize()
                    n_it = 9 if F == 30 else 3
                    ms_v, _ = median_ms(torch, lambda: score_video_device(model, frames, dl, scorer=scorer), n_it)
                aux[f"cfg3_4_3_30_F{F}_videos_per_s_per_gpu"] = 1e3 / ms_v
            # the same with 8 videos per backbone pass (score_videos_device; identical losses)
            vids = [synth_video(5 + i, 30).to(dev) for i in range(8)]
            with torch.no_grad():
                score_videos_device(model, vids, dl, scorer=scorer)
                torch.cuda.synchronize()
                ms_v, _ = median_ms(torch, lambda: score_videos_device(model, vids, dl, scorer=scorer), 3)
            aux["cfg3_4_3_30_F30_8_videos_per_pass_videos_per_s_per_gpu"] = 8e3 / ms_v
            del vids
        # MSVD-shaped, end to end, sharded over all ranks: mode H (bf16) and the reference's own mode R (precision "high")
        n_h = max(8, args.msvd_videos) * world
        aux["cfg3_msvd_modeH_bf16"] = cfg3_msvd_leg(torch, dist, model, dl, scorer, msvd_lengths(n_h), world, rank, dev)
        model_hi = VisionTransformer(img_size=224, num_classes=0, depth=12, qkv_bias=True,
                                     norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8, precision="high")
        model_hi.load_state_dict(sd, strict=True)
        model_hi = model_hi.eval().to(dev)
        n_r = max(4, args.msvd_videos // 3) * world
        aux["cfg3_msvd_modeR_high"] = cfg3_msvd_leg(torch, dist, model_hi, dl, None, msvd_lengths(n_r), world, rank, dev)
        if rank == 0:
            # the cfg2 step in precision="high" (2-term bf16 split in every GEMM and both attentions): what reference-exact
            # mode-R scoring costs next to the bf16 step
            xh = torch.randn(B_PER_GPU, 3, T, HW, HW, generator=torch.Generator().manual_seed(3)).to(dev)
            def step_hi():
                fh = model_hi(xh)
                return dl.pair_losses(fh, fh, s_idx, t_idx)
            with torch.no_grad():
                step_hi()
                torch.cuda.synchronize()
                aux["precision_high_ms_per_step"], _ = median_ms(torch, step_hi, 3)
            del xh
        aux["cfg3_msvd_note"] = ("MSVD-shaped = video lengths at evenly spaced quantiles of the 1 970 list lengths of the "
                                 "reference's loss_msvd_4_3_30.json; weak scaling (videos per GPU fixed), LPT sharding "
                                 "(parallel.shard_videos); e2e = uint8 frames from pinned host memory -> H2D -> "
                                 "dvst_preprocess_u8 -> window scoring -> gather_scores -> export_losses JSON -> "
                                 "dvst_select_frames; device = CUDA events around the rank's scoring, max over ranks; "
                                 "executed TFLOP/s by the de-duplicated window count of SURVEY.md section 8d")
        del model_hi
        torch.cuda.empty_cache()
        # ---- BASELINE configs[3]: frozen-backbone feature extraction for eval_knn (eval_knn.py:94-135): 256 clips per GPU
        # of the 8-frame global and the 4-frame local view shapes, CLS features + indices all-gathered to rank 0 ----
        from dvst_b200.parallel import extract_features
        cfg4 = {}
        for name, Tv in (("global_8f", 8), ("local_4f", 4)):
            xs = torch.randn(256, 3, Tv, HW, HW, generator=torch.Generator().manual_seed(7 + rank)).to(dev)
            idx = torch.arange(rank * 256, (rank + 1) * 256)
            batches = [(xs, idx)]
            extract_features(model, batches, 256 * world, dev)
            sync_all()
            ms4_local, feats = median_ms(torch, lambda: extract_features(model, batches, 256 * world, dev), 3, sync_all)
            tt = torch.tensor([ms4_local], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms4 = float(tt)
            cfg4[name] = {"clips_per_s": 256 * world / (ms4 / 1e3), "ms_per_batch": ms4,
                          "algorithmic_tflops_per_gpu": 256 * clip_gflop(Tv) / 1e3 / (ms4 / 1e3)}
            if rank == 0:
                assert feats.shape == (256 * world, 768) and bool(torch.isfinite(feats).all())
            del xs, feats
        cfg4["what"] = ("256 clips per GPU x (8-frame | 4-frame) x 224^2 through the frozen backbone, CLS features and sample "
                        "indices all-gathered to rank 0 (parallel.extract_features = eval_knn.py:94-135); device-timed incl. "
                        "the gather, max over ranks")
        aux["cfg4_knn_feature_extraction"] = cfg4
        torch.cuda.empty_cache()
        # ---- BASELINE configs[4]: long clips, 32 frames (T*N = 6 272 tokens per clip): divided vs joint space-time ----
        if rank == 0:
            cfg5 = {}
            for att, Bv in (("divided_space_time", 32), ("divided_space_time", 8), ("joint_space_time", 8)):
                if att == "divided_space_time":
                    m5 = model
                else:
                    m5 = VisionTransformer(img_size=224, num_classes=0, depth=12, qkv_bias=True,
                                           norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8,
                                           attention_type=att)
                    m5.load_state_dict(synth_backbone_state_dict(0, depth=12, num_frames=8, attention_type=att), strict=True)
                    m5 = m5.eval().to(dev)
                x5 = torch.randn(Bv, 3, 32, HW, HW, generator=torch.Generator().manual_seed(11)).to(dev)
                with torch.no_grad():
                    m5(x5)
                    torch.cuda.synchronize()
                    ms5, f5 = median_ms(torch, lambda: m5(x5), 5)
                gf = clip_gflop(32) if att == "divided_space_time" else 2523.6
                cfg5[f"{att}_B{Bv}"] = {"clips_per_s": Bv / (ms5 / 1e3), "ms_per_batch": ms5,
                                        "algorithmic_tflops": Bv * gf / 1e3 / (ms5 / 1e3)}
                assert bool(torch.isfinite(f5).all())
                del x5, f5
                if m5 is not model:
                    del m5
            cfg5["what"] = ("32 frames x 224^2 per clip (6 272 patch tokens + CLS), ViT-B depth 12, forward only, device-resident, "
                            "1 GPU: divided space-time (attention 3 % of the FLOPs) vs joint space-time (one 6 273-token "
                            "attention per clip and head, 57 % of the FLOPs, flash-style tcgen05 kernel); algorithmic FLOPs "
                            "of SURVEY.md section 8d")
            aux["cfg5_long_clip_32f"] = cfg5
            torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()

    rc = 0
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
            "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        # The GEMM kernel runs in two regimes: tensor-bound launches (qkv / fc1 consumers, fc2 producers, K|V, CLS-row GEMMs)
        # and the K = 768 deferred-LayerNorm producers (patch-embed, temporal proj, proj), which stream 1.85 GB per launch
        # for 0.24 TFLOP and are HBM-bound.  `roofline` is the dominant, tensor-bound class; the other is reported beside it.
        g_ms, g_flops, g_n = prof["gemm"]
        h_ms, h_bytes, h_n = prof["gemm_hbm"]
        a_ms, a_flops, a_n = prof["attention"]
        t_ms, t_flops, t_n = prof.get("gemm_tattn", (0.0, 0.0, 0))   # fused qkv GEMM + temporal attention launches
        achieved = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
        h_flops = h_n * 2.0 * (B_PER_GPU * 196 * T) * 768 * 768
        all_tf = (g_flops + h_flops + t_flops) / ((g_ms + h_ms + t_ms) / 1e3) / 1e12 if g_ms + h_ms + t_ms > 0 else 0.0
        traffic = traffic_hbm = traffic_fused = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
            traffic, traffic_hbm = tj["dram_bytes_per_launch"], tj.get("dram_bytes_per_launch_hbm_gemm")
            traffic_fused = tj.get("dram_bytes_per_launch_fused_temporal")
        except Exception:
            pass
        step_s = ms_max / args.steps / 1e3
        algo_tf = B_PER_GPU * GF_PER_CLIP_ALGO / 1e3 / step_s
        # executed = what the kernels were launched with (the library's own per-launch FLOP counters): GEMMs incl. the
        # folded temporal proj*fc and the last block's CLS-only half, attention 4 S^2 d per head
        exec_tf = (g_flops + h_flops + a_flops + t_flops) / args.steps / 1e12 / step_s
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel (tcgen05/TMEM), tensor-bound launches: qkv / fc1 "
                                                       "consumers, fc2 producers", "achieved": achieved,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                         "peak_source": peak_src, "launches": g_n, "ms_in_kernel_per_step": g_ms / args.steps,
                         "flops_per_launch_avg": g_flops / max(g_n, 1),
                         "all_gemm_launches_tflops": all_tf},
            "roofline_hbm_gemm": {"bound": "hbm", "kernel": "gemm_bf16_tn_kernel<2,3,0,0> with K = 768 (deferred-LayerNorm "
                                  "producers: residual in, fp32 + bf16 out)", "achieved": h_bytes / (h_ms / 1e3) / 1e9 if h_ms > 0 else 0.0,
                                  "peak": peaks.get("hbm_gbs", 6558.4), "unit": "GB/s",
                                  "frac": (h_bytes / (h_ms / 1e3) / 1e9 / peaks.get("hbm_gbs", 6558.4)) if h_ms > 0 else 0.0,
                                  "traffic": traffic_hbm, "launches": h_n, "ms_in_kernel_per_step": h_ms / args.steps,
                                  "bytes_per_launch_avg": h_bytes / max(h_n, 1)},
            # the temporal half of every block: deferred-LayerNorm qkv GEMM whose epilogue runs the temporal attention out of
            # shared memory (gemm_ln_tattn_kernel; FLOPs = the qkv GEMM + 4 T^2 64 per sequence and head)
            "roofline_fused_temporal": {"bound": "tensor", "kernel": "gemm_ln_tattn_kernel (tcgen05 256x192 head tiles + "
                                        "mma.sync temporal attention in the epilogue; no packed qkv tensor in HBM)",
                                        "achieved": t_flops / (t_ms / 1e3) / 1e12 if t_ms > 0 else 0.0, "peak": peak_tf,
                                        "unit": "TFLOP/s", "frac": (t_flops / (t_ms / 1e3) / 1e12 / peak_tf) if t_ms > 0 else 0.0,
                                        "traffic": traffic_fused, "launches": t_n, "ms_in_kernel_per_step": t_ms / args.steps},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            # the step against the bf16 peak on EXECUTED FLOPs (never on skipped work); the reference-formulation figure
            # (784.1 GF per clip, counts the un-folded temporal_fc and the last block's dead patch-token work) beside it
            "step_executed_tflops_per_gpu": exec_tf,
            "step_frac_of_bf16_burst_peak": exec_tf / peaks.get("bf16_tflops", 1642.8),
            "step_frac_of_bf16_sustained_peak": exec_tf / peak_tf,
            "step_algorithmic_tflops_per_gpu_reference_formulation": algo_tf,
        }
        if aux:
            line["aux"] = aux
        if world == 1:
            arm, vps, s_per_step, f_cpu, l_cpu = run_cpu_arm(8, 1, 2)
            line["cpu_baseline"] = {"value": vps, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
                                    "sample": arm.describe(2, 8) + " (after 1 warm-up)"}
            fp, lp, same_bits = gpu_par
            feat_rel = float((fp.double() - f_cpu.double()).norm() / f_cpu.double().norm())
            loss_rel = float(((lp.double() - l_cpu.double()).abs() / l_cpu.double().abs()).max())
            ok = feat_rel <= PARITY_FEAT_TOL and loss_rel <= PARITY_LOSS_TOL and same_bits
            line["parity"] = {"feat_rel": feat_rel, "loss_rel": loss_rel, "feat_tol": PARITY_FEAT_TOL,
                              "loss_tol": PARITY_LOSS_TOL, "checker": arm.kind,
                              "bits_equal_inside_64_clip_batch": same_bits, "ok": ok,
                              "what": "two cfg2 clips (seed 1): GPU backbone features and mode-H pair loss vs the CPU arm"}
            if not ok:
                rc = 1
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rc:
        print("[bench] PARITY FAILED - see the `parity` key", file=sys.stderr)
        sys.exit(rc)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
