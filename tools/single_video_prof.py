"""Where does the time of scoring ONE short video (F = 30: 28 three-frame student clips + one 30-frame teacher clip) go?
Host time per call (perf_counter, no sync) vs device time (CUDA events) of each piece, eager vs CUDA-graph replay."""
import os, sys, time, torch
from functools import partial
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200
from dvst_b200 import _lib
from dvst_b200.dino_similarity import DINOLoss, HeadScorer, _forward_windows, plan_windows, score_video_device
from dvst_b200.models.timesformer import VisionTransformer
from dvst_b200.testing import synth_backbone_state_dict, synth_head_state_dict, synth_video
from dvst_b200.vision_transformer import DINOHead
dev = torch.device("cuda:0")
m = VisionTransformer(img_size=224, num_classes=0, depth=12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
m.load_state_dict(synth_backbone_state_dict(0, depth=12, num_frames=8), strict=True)
m = m.eval().to(dev)
head = DINOHead(768, 65536); head.load_state_dict(synth_head_state_dict(2), strict=True); head = head.to(dev).eval()
scorer = HeadScorer(head, 0.04, 0.1)
dl = DINOLoss(768, 0.02, 0.3).to(dev)
frames = synth_video(5, 30).to(dev)


def timeit(name, fn, n=5):
    with torch.no_grad():
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.launch_count()
        t0 = time.perf_counter(); e0.record()
        for _ in range(n): fn()
        e1.record(); host = (time.perf_counter() - t0) / n
        torch.cuda.synchronize()
        print(f"{name:46s} host {host * 1e3:7.3f} ms   device {e0.elapsed_time(e1) / n:7.3f} ms   launches {(_lib.launch_count() - n0) // n}", flush=True)


s_starts, t_starts, s_idx, t_idx, groups = plan_windows([30], 3, 30, dev)
timeit("plan_windows", lambda: plan_windows([30], 3, 30, dev))
timeit("student forward (28 clips x T=3)", lambda: _forward_windows(m, frames, s_starts, 3, 64))
timeit("teacher forward (1 clip x T=30)", lambda: _forward_windows(m, frames, t_starts, 30, 64))
fs = _forward_windows(m, frames, s_starts, 3, 64); ft = _forward_windows(m, frames, t_starts, 30, 64)
timeit("mode-H loss (bottleneck + fused head CE)", lambda: scorer.pair_losses(fs, ft, s_idx, t_idx))
timeit("mode-R loss", lambda: dl.pair_losses(fs, ft, s_idx, t_idx))
timeit("score_video_device mode H (whole video)", lambda: score_video_device(m, frames, dl, scorer=scorer))
timeit("score_video_device mode R (whole video)", lambda: score_video_device(m, frames, dl))
timeit("  same, single stream", lambda: score_video_device(m, frames, dl, small_batch_rows=0))
# CUDA graph of the whole mode-R scoring of this video
with torch.no_grad():
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        score_video_device(m, frames, dl, small_batch_rows=0)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = score_video_device(m, frames, dl, small_batch_rows=0)
    ref = score_video_device(m, frames, dl, small_batch_rows=0)
    g.replay(); torch.cuda.synchronize()
    print("graph replay equals eager:", bool(torch.equal(out, ref)))
timeit("CUDA-graph replay, mode R, single stream", lambda: g.replay())
