"""Small-batch forward latency: eager launches vs a captured CUDA graph (the forward is capturable: no sync / allocation
inside the C ABI).  python tools/graph_latency.py"""
import os, sys, time, torch
from functools import partial
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dvst_b200  # noqa: F401
from dvst_b200.models.timesformer import VisionTransformer
from dvst_b200.testing import synth_backbone_state_dict, synth_clips
dev = torch.device("cuda:0")
m = VisionTransformer(img_size=224, num_classes=0, depth=12, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_frames=8)
m.load_state_dict(synth_backbone_state_dict(0, depth=12, num_frames=8), strict=True)
m = m.eval().to(dev)
for B, T in ((1, 3), (1, 8), (1, 16), (2, 8), (1, 30), (8, 16)):
    x = synth_clips(1, B, T).to(dev)
    with torch.no_grad():
        for _ in range(3): m(x)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): m(x)
        torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / 20
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side): m(x)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): out = m(x)
        for _ in range(3): g.replay()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): g.replay()
        torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / 20
    print(f"B={B} T={T}: eager {eager * 1e3:.3f} ms, graph replay {graph * 1e3:.3f} ms", flush=True)
