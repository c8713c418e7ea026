"""CPU oracle for the dvst_b200 hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-the-specification restatement (plain torch CPU tensor arithmetic, no nn.Module, functional, its
own [B,N,T,D] + separate-CLS token layout) of the reference's TimeSformer forward, DINO head, DINO losses, scoring
window rule and adaptive frame selection.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package (`dvst_b200`) never does.

Parity pinning: the reference ships no tests / golden vectors for this path, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, imported in the authoring container through `oracle/ref_shim.py`:
`oracle/make_golden.py` writes `tests/golden/*.pt` (inputs are regenerated from seeds, outputs are stored), and
`tests/test_oracle_golden.py` checks this restatement against those fixtures on every CPU run.

Each function cites the reference lines it restates (paths relative to the reference repo root).
"""
import math

import numpy as np
import torch

PATCH = 16
EMBED = 768
HEADS = 12
HEAD_DIM = 64
LN_EPS = 1e-6


# ---------------------------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------------------------
def layer_norm(x, w, b):
    """nn.LayerNorm(eps=1e-6) as built by partial(nn.LayerNorm, eps=1e-6) - models/timesformer.py:596."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def gelu(x):
    """nn.GELU() default = exact erf form - models/timesformer.py:37,42."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(x, w, b=None):
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def nearest_index(dst_len, src_len):
    """Index map of F.interpolate(mode='nearest') used for time_embed (models/timesformer.py:314-318) and pos_embed
    (:292-303): src = min(floor(dst * float32(src_len / dst_len)), src_len - 1), computed in float32 like ATen."""
    scale = np.float32(src_len) / np.float32(dst_len)
    idx = np.floor(np.arange(dst_len, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, src_len - 1)


def mha(q, k, v):
    """softmax(q k^T / sqrt(64)) v per head - models/timesformer.py:77-81.  q,k,v: [..., S, 768]."""
    S = q.shape[-2]
    lead = q.shape[:-2]
    qh = q.reshape(*lead, S, HEADS, HEAD_DIM).transpose(-2, -3)
    kh = k.reshape(*lead, k.shape[-2], HEADS, HEAD_DIM).transpose(-2, -3)
    vh = v.reshape(*lead, v.shape[-2], HEADS, HEAD_DIM).transpose(-2, -3)
    s = (qh @ kh.transpose(-1, -2)) * (HEAD_DIM ** -0.5)
    p = torch.softmax(s, dim=-1)
    o = (p @ vh).transpose(-2, -3)
    return o.reshape(*lead, S, HEADS * HEAD_DIM)


# ---------------------------------------------------------------------------------------------------------------------
# TimeSformer forward (divided space-time, joint space-time, space-only)
# ---------------------------------------------------------------------------------------------------------------------
def embed_tokens(sd, x, attention_type="divided_space_time"):
    """PatchEmbed + cls/pos/time embedding assembly - models/timesformer.py:203-209,287-323.

    Returns tok [B,N,T,D] (patch tokens, n-major / t-minor like the reference's `b (n t) m`), cls [B,D]."""
    B, C, T, H, W = x.shape
    gh, gw = H // PATCH, W // PATCH
    N = gh * gw
    patches = x.reshape(B, C, T, gh, PATCH, gw, PATCH).permute(0, 2, 3, 5, 1, 4, 6).reshape(B, T, N, C * PATCH * PATCH)
    wpe = sd["patch_embed.proj.weight"].reshape(EMBED, -1)
    tok = linear(patches, wpe, sd["patch_embed.proj.bias"])  # [B,T,N,D]
    pos = sd["pos_embed"][0]  # [1+P*P, D]
    if N + 1 != pos.shape[0]:
        P = int(round(math.sqrt(pos.shape[0] - 1)))
        rows = nearest_index((N + 1) // gw, P)  # reference derives the grid height as (N+1)//W, :297
        cols = nearest_index(gw, P)
        grid = (rows[:, None] * P + cols[None, :]).reshape(-1)
        pos_patch = pos[1:][torch.from_numpy(grid)]
    else:
        pos_patch = pos[1:]
    tok = tok + pos_patch  # broadcast over (B,T)
    cls = (sd["cls_token"][0, 0] + pos[0]).expand(B, EMBED).clone()
    if attention_type != "space_only":
        te = sd["time_embed"][0]  # [T0, D]
        if T != te.shape[0]:
            te = te[torch.from_numpy(nearest_index(T, te.shape[0]))]
        tok = tok + te[None, :, None, :]
    return tok.permute(0, 2, 1, 3).contiguous(), cls


def block_divided(sd, pre, tok, cls):
    """Block.forward, attention_type == 'divided_space_time' - models/timesformer.py:124-185."""
    B, N, T, D = tok.shape
    g = lambda k: sd[pre + k]
    # temporal attention over the T tokens of each (b, n); CLS excluded (:127-135)
    u = layer_norm(tok, g("temporal_norm1.weight"), g("temporal_norm1.bias"))
    q, k, v = linear(u, g("temporal_attn.qkv.weight"), g("temporal_attn.qkv.bias")).split(D, dim=-1)
    a = mha(q, k, v)  # sequence axis = T
    r = linear(a, g("temporal_attn.proj.weight"), g("temporal_attn.proj.bias"))
    r = linear(r, g("temporal_fc.weight"), g("temporal_fc.bias"))
    tok = tok + r
    # spatial attention per (b, t) over [cls, N patches] (:140-159)
    seq = torch.cat([cls[:, None, None, :].expand(B, T, 1, D), tok.permute(0, 2, 1, 3)], dim=2)  # [B,T,1+N,D]
    u = layer_norm(seq, g("norm1.weight"), g("norm1.bias"))
    q, k, v = linear(u, g("attn.qkv.weight"), g("attn.qkv.bias")).split(D, dim=-1)
    a = mha(q, k, v)
    r = linear(a, g("attn.proj.weight"), g("attn.proj.bias"))  # [B,T,1+N,D]
    cls = cls + r[:, :, 0].mean(dim=1)  # CLS residual averaged over frames (:162-164)
    tok = tok + r[:, :, 1:].permute(0, 2, 1, 3)
    # MLP on every row (:184)
    tok = tok + mlp(sd, pre, layer_norm(tok, g("norm2.weight"), g("norm2.bias")))
    cls = cls + mlp(sd, pre, layer_norm(cls, g("norm2.weight"), g("norm2.bias")))
    return tok, cls


def mlp(sd, pre, u):
    """Mlp.forward - models/timesformer.py:46-52."""
    h = gelu(linear(u, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
    return linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])


def block_plain(sd, pre, seq):
    """Block.forward for 'space_only' / 'joint_space_time' - models/timesformer.py:120-123.  seq [..., S, D]."""
    g = lambda k: sd[pre + k]
    u = layer_norm(seq, g("norm1.weight"), g("norm1.bias"))
    q, k, v = linear(u, g("attn.qkv.weight"), g("attn.qkv.bias")).split(EMBED, dim=-1)
    seq = seq + linear(mha(q, k, v), g("attn.proj.weight"), g("attn.proj.bias"))
    return seq + mlp(sd, pre, layer_norm(seq, g("norm2.weight"), g("norm2.bias")))


def depth_of(sd):
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def forward_features(sd, x, attention_type="divided_space_time", get_all=False):
    """VisionTransformer.forward_features / forward(x, use_head=False) - models/timesformer.py:285-351.

    sd: state dict with the reference's key names; x [B,3,T,H,W].  Returns the LayerNorm'ed CLS feature [B,768]
    (or all tokens [B,1+N*T,768] in the reference's order if get_all)."""
    x = x.to(sd["pos_embed"].dtype)
    tok, cls = embed_tokens(sd, x, attention_type)
    B, N, T, D = tok.shape
    depth = depth_of(sd)
    if attention_type == "divided_space_time":
        for i in range(depth):
            tok, cls = block_divided(sd, f"blocks.{i}.", tok, cls)
        allt = torch.cat([cls[:, None], tok.reshape(B, N * T, D)], dim=1)
    elif attention_type == "joint_space_time":
        seq = torch.cat([cls[:, None], tok.reshape(B, N * T, D)], dim=1)
        for i in range(depth):
            seq = block_plain(sd, f"blocks.{i}.", seq)
        allt = seq
    elif attention_type == "space_only":
        seq = torch.cat([cls[:, None, None].expand(B, T, 1, D), tok.permute(0, 2, 1, 3)], dim=2)  # [B,T,1+N,D]
        for i in range(depth):
            seq = block_plain(sd, f"blocks.{i}.", seq)
        allt = seq.mean(dim=1)  # frame average (:338-340)
    else:
        raise ValueError(attention_type)
    out = layer_norm(allt, sd["norm.weight"], sd["norm.bias"])
    return out if get_all else out[:, 0]


def attention_map(sd, x):
    """forward_features(x, get_attn=True), divided_space_time - models/timesformer.py:325-331 with Block.forward's
    return_attn branch (:155-157): blocks 0..depth-2 run in full, the last block applies its temporal half, then returns
    softmax(q k^T / 8) of the spatial attention over [cls, N patches] of every frame: [B*T, 12, 1+N, 1+N]."""
    x = x.to(sd["pos_embed"].dtype)
    tok, cls = embed_tokens(sd, x, "divided_space_time")
    B, N, T, D = tok.shape
    depth = depth_of(sd)
    for i in range(depth - 1):
        tok, cls = block_divided(sd, f"blocks.{i}.", tok, cls)
    pre = f"blocks.{depth - 1}."
    g = lambda k: sd[pre + k]
    u = layer_norm(tok, g("temporal_norm1.weight"), g("temporal_norm1.bias"))
    q, k, v = linear(u, g("temporal_attn.qkv.weight"), g("temporal_attn.qkv.bias")).split(D, dim=-1)
    r = linear(mha(q, k, v), g("temporal_attn.proj.weight"), g("temporal_attn.proj.bias"))
    tok = tok + linear(r, g("temporal_fc.weight"), g("temporal_fc.bias"))
    seq = torch.cat([cls[:, None, None, :].expand(B, T, 1, D), tok.permute(0, 2, 1, 3)], dim=2)  # [B,T,1+N,D]
    u = layer_norm(seq, g("norm1.weight"), g("norm1.bias"))
    q, k, _ = linear(u, g("attn.qkv.weight"), g("attn.qkv.bias")).split(D, dim=-1)
    qh = q.reshape(B * T, N + 1, HEADS, HEAD_DIM).transpose(1, 2)
    kh = k.reshape(B * T, N + 1, HEADS, HEAD_DIM).transpose(1, 2)
    return torch.softmax((qh @ kh.transpose(-1, -2)) * (HEAD_DIM ** -0.5), dim=-1)


# ---------------------------------------------------------------------------------------------------------------------
# DINO head and losses
# ---------------------------------------------------------------------------------------------------------------------
def dino_head(hd, x):
    """DINOHead.forward (nlayers=3, no BN) - vision_transformer.py:248-282.  hd keys: mlp.{0,2,4}.{weight,bias},
    last_layer.weight_g [out,1], last_layer.weight_v [out,256] (deprecated weight_norm hook, dim=0)."""
    h = gelu(linear(x, hd["mlp.0.weight"], hd["mlp.0.bias"]))
    h = gelu(linear(h, hd["mlp.2.weight"], hd["mlp.2.bias"]))
    z = linear(h, hd["mlp.4.weight"], hd["mlp.4.bias"])
    z = z / z.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    v = hd["last_layer.weight_v"]
    w = hd["last_layer.weight_g"] * v / v.norm(dim=1, keepdim=True)
    return linear(z, w)


def dino_pair_loss(student, teacher, center, teacher_temp, student_temp):
    """Scoring DINOLoss.forward - dino_similarity.py:129-135.  student/teacher [..., out_dim] -> [...] losses
    (the reference calls it on single vectors; .mean() of a 1-element sum is the value itself)."""
    t = torch.softmax((teacher - center) / teacher_temp, dim=-1)
    ls = torch.log_softmax(student / student_temp, dim=-1)
    return -(t * ls).sum(dim=-1)


def dino_train_loss(student_out, teacher_out, center, teacher_temp, student_temp, ncrops, global_crops=2):
    """Training DINOLoss.forward (two_token=False) - train_ssl.py:653-668 (center update excluded, see below)."""
    s = (student_out / student_temp).chunk(ncrops)
    t = torch.softmax((teacher_out - center) / teacher_temp, dim=-1).chunk(global_crops)
    total, n = 0.0, 0
    for iq, q in enumerate(t):
        for v in range(len(s)):
            if v == iq:
                continue
            total = total + (-(q * torch.log_softmax(s[v], dim=-1)).sum(dim=-1)).mean()
            n += 1
    return total / n


def center_update(center, teacher_out_per_rank, momentum=0.9):
    """DINOLoss.update_center - train_ssl.py:686-691.  teacher_out_per_rank: list (one per rank) of [B,out_dim]."""
    world = len(teacher_out_per_rank)
    batch_center = sum(t.sum(dim=0, keepdim=True) for t in teacher_out_per_rank)
    batch_center = batch_center / (len(teacher_out_per_rank[0]) * world)
    return center * momentum + batch_center * (1 - momentum)


# ---------------------------------------------------------------------------------------------------------------------
# scoring windows, whole-video scoring, selection
# ---------------------------------------------------------------------------------------------------------------------
def window_starts(F, local_size=3, global_size=30):
    """get_views_of_video_same_size - datasets_custom/dino_loss_loader.py:126-171.
    Returns (G, [(j, l)] per frame): student clip = frames[j:j+local_size], teacher clip = frames[l:l+G]."""
    loc, glob = local_size // 2, global_size // 2
    G = min(global_size, F)
    out = []
    for i in range(F):
        j, k, l, m = i - loc, i + loc + 1, i - glob, i + glob
        if j < 0:
            j, k = 0, local_size
        if k >= F:
            k, j = F, F - local_size
        if l < 0:
            l, m = 0, G
        if m >= F:
            m, l = F, F - G
        out.append((j, l))
    return G, out


def preprocess_frames(video_uint8, sampling_rate=4, mean=(0.45, 0.45, 0.45), std=(0.225, 0.225, 0.225), crop_size=224):
    """Loader front half - datasets_custom/dino_loss_loader.py:80-92: every `sampling_rate`-th decoded frame, / 255,
    - mean, / std as three separately rounded fp32 operations (data_utils.py:308-325), T H W C -> T C H W, centre crop
    with ceil((H - crop) / 2) offsets (transform.py:206-249, spatial_idx = 1)."""
    frames = video_uint8[::sampling_rate].to(torch.float32) / 255.0
    frames = frames - torch.tensor(mean, dtype=torch.float32)
    frames = frames / torch.tensor(std, dtype=torch.float32)
    frames = frames.permute(0, 3, 1, 2)
    h, w = frames.shape[2], frames.shape[3]
    y0, x0 = int(math.ceil((h - crop_size) / 2)), int(math.ceil((w - crop_size) / 2))
    return frames[:, :, y0:y0 + crop_size, x0:x0 + crop_size].contiguous()


def short_video_clips(F, local_size=3):
    """Frame indices of the (student, teacher) clips every frame of a video with F < local_size sampled frames is scored
    on (dino_loss_loader.py:126-171 with len(frames) < local_size, then dino_similarity.py:78): G = F; the local slice
    frames[F - local_size : F] starts at a negative index, i.e. at max(2F - local_size, 0); its n frames are written into
    the zero [3, G, H, W] tensor through `tensor_padded[:, :local_size] = tensor_local`, which broadcasts n = 1 to all G
    frames, fits n = G exactly, and raises for anything else (as the reference does).  The teacher clip is frames[0:F]."""
    start = max(2 * F - local_size, 0)
    n = F - start
    if n == 1:
        s_idx = [start] * F
    elif n == F:
        s_idx = list(range(F))
    else:
        raise ValueError(f"the reference cannot build views for {F} frames with local_clip_size {local_size} "
                         f"(broadcast of {n} frames into {F})")
    return s_idx, list(range(F))


def score_video(sd, frames, local_size=3, global_size=30, teacher_temp=0.02, student_temp=0.3, head=None,
                center=None, attention_type="divided_space_time", batch=8):
    """dino_similarity() inner loop - dino_similarity.py:66-93 (mode R when head is None; mode H = DINOHead on both
    features first).  frames [F,3,H,W] (already normalised / cropped).  Returns a python list of F floats.
    Wrong crops follow the dummy path (dino_loss_loader.py:36-38,103-107): `global_size` losses of the all-zero
    (T=local_size, T=60) pair.  Videos shorter than the local window (F < local_size) are NOT dummies in the reference:
    get_views_of_video_same_size clamps G = F, its local slice frames[F-local_size:F] has a negative start (python
    wrap-around), and the zero-padded [3,G,H,W] tensor receives it by broadcast (dino_loss_loader.py:147-161), so the
    student sees T = F frames; every frame of the video gets that same pair (short_video_clips)."""
    F = frames.shape[0]
    dt = sd["pos_embed"].dtype
    if frames.shape[-1] != 224 or frames.shape[-2] != 224 or F == 0:
        zs = torch.zeros(1, 3, local_size, 224, 224, dtype=dt)
        zt = torch.zeros(1, 3, 60, 224, 224, dtype=dt)
        fs, ft = forward_features(sd, zs, attention_type), forward_features(sd, zt, attention_type)
        if head is not None:
            fs, ft = dino_head(head, fs), dino_head(head, ft)
        c = torch.zeros_like(fs) if center is None else center
        return [float(dino_pair_loss(fs, ft, c, teacher_temp, student_temp)[0])] * global_size
    if F < local_size:
        s_idx, t_idx = short_video_clips(F, local_size)
        fs = forward_features(sd, frames[s_idx].permute(1, 0, 2, 3)[None].to(dt), attention_type)
        ft = forward_features(sd, frames[t_idx].permute(1, 0, 2, 3)[None].to(dt), attention_type)
        if head is not None:
            fs, ft = dino_head(head, fs), dino_head(head, ft)
        c = torch.zeros_like(fs) if center is None else center
        return [float(dino_pair_loss(fs, ft, c, teacher_temp, student_temp)[0])] * F
    G, starts = window_starts(F, local_size, global_size)
    cache = {}

    def feat(start, length):
        key = (start, length)
        if key not in cache:
            clip = frames[start:start + length].permute(1, 0, 2, 3)[None].to(dt)
            f = forward_features(sd, clip, attention_type)
            cache[key] = dino_head(head, f) if head is not None else f
        return cache[key]

    losses = []
    for (j, l) in starts:
        s, t = feat(j, local_size), feat(l, G)
        c = torch.zeros_like(s) if center is None else center
        losses.append(float(dino_pair_loss(s, t, c, teacher_temp, student_temp)[0]))
    return losses


def select_frames(loss_list, n_unsampled, rate, N, n_sampled=None):
    """Adaptive selection - datasets_custom/frame_selection_loader.py:115-190 (inverse-CDF quantile sampling).
    loss_list: per-sampled-frame losses; n_unsampled = number of frames of the original video; n_sampled = number of
    pre-sampled frames (defaults to ceil(n_unsampled / rate)).  Returns N python ints."""
    if n_sampled is None:
        n_sampled = -(-n_unsampled // rate)
    loss = np.asarray(loss_list, dtype=np.float64)[:n_sampled]
    cdf = np.cumsum(loss / np.sum(loss))
    picked = []
    for i in range(N):
        idx = int(np.abs(cdf - i / N).argmin())
        s = idx * rate
        if s not in picked:
            picked.append(s)
            continue
        s += rate
        while True:
            if s not in picked and s < n_unsampled:
                picked.append(s)
                break
            if s >= n_unsampled:
                picked.append(s - rate)
                break
            s += rate
    return picked


def knn_classifier(train_features, train_labels, test_features, test_labels, k, T, num_classes=1000):
    """Weighted k-NN vote - eval_knn.py:138-178.  similarity = test @ train.T (:153), top-k neighbours (:154; equal
    similarities by ascending index), per-class sums of exp(sim / T) over the neighbours in order (:155-166), then the
    position of the true class in the descending sort of the class scores (:167-172; equal scores by ascending class
    index, i.e. a stable sort).  Returns (top1 %, top5 %, rank [n_test] int64, neighbour indices [n_test, k] int64);
    the reference returns the first two."""
    train = train_features.to(torch.float32)
    test = test_features.to(torch.float32)
    sim = test @ train.t()
    order = torch.sort(sim, dim=1, descending=True, stable=True)
    dist, idx = order.values[:, :k], order.indices[:, :k]
    w = (dist / T).exp()
    labels = train_labels.view(-1)[idx]
    n = test.shape[0]
    probs = torch.zeros(n, num_classes, dtype=torch.float32)
    for j in range(k):  # sequential over neighbours like torch.sum over dim 1 of the [B, k, C] product
        probs[torch.arange(n), labels[:, j]] += w[:, j]
    tgt = test_labels.view(-1)
    pt = probs[torch.arange(n), tgt].unsqueeze(1)
    cls = torch.arange(num_classes).unsqueeze(0)
    rank = ((probs > pt) | ((probs == pt) & (cls < tgt.unsqueeze(1)))).sum(1)
    top1 = float((rank == 0).sum()) * 100.0 / n
    top5 = float((rank < 5).sum()) * 100.0 / n
    return top1, top5, rank, idx
