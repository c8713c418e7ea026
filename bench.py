#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native scoring path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Metric (BASELINE.json): videos/sec scored.  Workload = BASELINE.json configs[1]: TimeSformer ViT-B/16, 16 frames,
224^2, Kinetics-400-shaped synthetic clips, batch 64 per GPU.  One step = 64 clips (1 clip = 1 video) through the
backbone forward, the DINO projection head and the fused centred / sharpened cross-entropy on the 32 (student, teacher)
pairs formed by (clip 2i, clip 2i+1); weak scaling (every rank scores its own 64 videos, no data-path collective; the
32 per-rank losses are all-gathered).  `value` is timed with inputs resident in HBM; `e2e` goes through the public
module API from pinned HOST buffers (H2D of the step's clips and D2H of its losses inside the timed region).
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
# algorithmic FLOPs per clip in the reference formulation, T = 16, N = 196
GF_PER_CLIP_ALGO = 784.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-aux", action="store_true", help="skip the auxiliary config-3 (4_3_30 video) numbers")
    return ap.parse_args()


def workload_config(n_gpus):
    return {"workload": "cfg2: TimeSformer ViT-B/16 divided space-time, 64 clips x 16 frames x 224^2 per GPU per step "
                        "(Kinetics-400-shaped, 1 clip = 1 video): backbone forward + DINOHead(768->65536) + fused "
                        "centred/sharpened CE on 32 (student, teacher) pairs",
            "clips_per_gpu": B_PER_GPU, "frames": T, "crop": HW, "pairs_per_gpu": B_PER_GPU // 2, "out_dim": OUT_DIM,
            "teacher_temp": TT, "student_temp": TS, "weights": "random init, seed 0 (all tensors randomised)",
            "parallelism": f"dp{n_gpus} (videos sharded, replicated frozen weights)",
            "l2_policy": "inputs exceed L2 (616 MB of clips per step vs 126 MB L2); activations 3 GB per step",
            "precision": "bf16 operands, fp32 accumulate / residual / LayerNorm / softmax"}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's PyTorch path on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_step(sd, hsd, clips, O, torch):
    """Same step on `clips` [B,3,T,224,224] (B even) with the CPU oracle."""
    with torch.no_grad():
        f = O.forward_features(sd, clips)
        logits = O.dino_head(hsd, f)
        loss = O.dino_pair_loss(logits[0::2], logits[1::2], torch.zeros(1, OUT_DIM), TT, TS)
    return loss


def run_cpu_port(n_steps, n_warm, clips_per_step):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dvst_oracle as O
    import dvst_b200  # noqa: F401
    from dvst_b200.testing import synth_backbone_state_dict, synth_clips, synth_head_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth_backbone_state_dict(0, depth=12, num_frames=8)
    hsd = synth_head_state_dict(2, out_dim=OUT_DIM)
    clips = synth_clips(1, clips_per_step, T)
    for _ in range(n_warm):
        cpu_port_step(sd, hsd, clips, O, torch)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        cpu_port_step(sd, hsd, clips, O, torch)
    dt = time.perf_counter() - t0
    return clips_per_step * n_steps / dt, dt / n_steps, torch.get_num_threads()


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips = 2
    steps, warm = max(1, min(args.steps, 4)), max(1, min(args.warmup, 1))
    vps, s_per_step, cores = run_cpu_port(steps, warm, clips)
    line = {"impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": vps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{clips} clips (1 pair) of the cfg2 step per timed step, {steps} steps; "
                                       "oracle/dvst_oracle.py = CPU restatement of the reference's PyTorch path "
                                       "(the reference itself does not travel to the GPU box)"},
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
    from dvst_b200.dino_similarity import DINOLoss, HeadScorer, score_video_device, score_videos_device
    from dvst_b200.models.timesformer import VisionTransformer
    from dvst_b200.testing import synth_backbone_state_dict, synth_head_state_dict, synth_video
    from dvst_b200.vision_transformer import DINOHead
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    model = VisionTransformer(img_size=224, num_classes=0, depth=12, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
    model.load_state_dict(synth_backbone_state_dict(0, depth=12, num_frames=8), strict=True)
    model = model.eval().to(dev)
    head = DINOHead(768, OUT_DIM)
    head.load_state_dict(synth_head_state_dict(2, out_dim=OUT_DIM), strict=True)
    head = head.to(dev).eval()
    scorer = HeadScorer(head, TT, TS)
    P = B_PER_GPU // 2
    s_idx = torch.arange(0, B_PER_GPU, 2, dtype=torch.int32, device=dev)
    t_idx = torch.arange(1, B_PER_GPU, 2, dtype=torch.int32, device=dev)

    g = torch.Generator().manual_seed(1234 + rank)
    host_clips = [torch.randn(B_PER_GPU, 3, T, HW, HW, generator=g).pin_memory() for _ in range(2)]
    dev_clips = [h.to(dev) for h in host_clips]
    gathered = torch.empty(world * P, dtype=torch.float32, device=dev) if world > 1 else None

    def step(x):
        with torch.no_grad():
            f = model(x)
            loss = scorer.pair_losses(f, f, s_idx, t_idx)
            if world > 1:
                dist.all_gather_into_tensor(gathered, loss)
                return gathered
            return loss

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

    # ---- end to end through the public API from pinned host buffers (double-buffered upload, deferred readback) ----
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty_like(dev_clips[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    host_loss = [torch.empty(world * P if world > 1 else P, dtype=torch.float32).pin_memory() for _ in range(2)]

    def e2e_run(n_steps):
        results = []
        cur = torch.cuda.current_stream()
        pending = None
        for f_ in free:
            f_.record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[0])
            stage[0].copy_(host_clips[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n_steps):
            b = i % 2
            if i + 1 < n_steps:
                nb = (i + 1) % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[nb])
                    stage[nb].copy_(host_clips[nb], non_blocking=True)
                    ready[nb].record(copy_stream)
            cur.wait_event(ready[b])
            loss = step(stage[b])
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
    assert len(res) == args.steps
    t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B_PER_GPU * args.steps / float(t_e.item())
    h2d = host_clips[0].numel() * 4
    d2h = host_loss[0].numel() * 4

    # ---- auxiliary: config 3 (4_3_30 adaptive-frame-selection scoring of whole videos), device resident ----
    aux = None
    if rank == 0 and not args.no_aux:
        aux = {}
        dl = DINOLoss(768, 0.02, 0.3).to(dev)
        for F in (30, 68):
            frames = synth_video(5, F).to(dev)
            with torch.no_grad():
                score_video_device(model, frames, dl)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(3):
                    score_video_device(model, frames, dl, scorer=scorer)
                a1.record()
                torch.cuda.synchronize()
            aux[f"cfg3_4_3_30_F{F}_videos_per_s_per_gpu"] = 3e3 / a0.elapsed_time(a1)
        # the same with 8 videos per backbone pass (score_videos_device; identical losses)
        vids = [synth_video(5 + i, 30).to(dev) for i in range(8)]
        with torch.no_grad():
            score_videos_device(model, vids, dl, scorer=scorer)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(3):
                score_videos_device(model, vids, dl, scorer=scorer)
            a1.record()
            torch.cuda.synchronize()
        aux["cfg3_4_3_30_F30_8_videos_per_pass_videos_per_s_per_gpu"] = 24e3 / a0.elapsed_time(a1)
    if world > 1:
        dist.barrier()

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
        achieved = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
        h_flops = h_n * 2.0 * (B_PER_GPU * 196 * T) * 768 * 768
        all_tf = (g_flops + h_flops) / ((g_ms + h_ms) / 1e3) / 1e12 if g_ms + h_ms > 0 else 0.0
        traffic = traffic_hbm = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
            traffic, traffic_hbm = tj["dram_bytes_per_launch"], tj.get("dram_bytes_per_launch_hbm_gemm")
        except Exception:
            pass
        step_tf = B_PER_GPU * GF_PER_CLIP_ALGO / 1e3 / (ms_max / args.steps / 1e3)
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
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "step_algorithmic_tflops_per_gpu": step_tf,
            "step_frac_of_bf16_burst_peak": step_tf / peaks.get("bf16_tflops", 1642.8),
        }
        if aux:
            line["aux"] = aux
        if world == 1:
            vps, s_per_step, cores = run_cpu_port(2, 1, 2)
            line["cpu_baseline"] = {"value": vps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "2 clips (1 pair) of the cfg2 step, 2 timed steps after 1 warm-up; "
                                              "oracle/dvst_oracle.py (torch CPU fp32, all host threads)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
