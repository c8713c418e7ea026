"""Kernel-level parity through the C ABI on a B200 (each kernel vs a plain fp32 torch statement of the same op)."""
import pytest
import torch

import dvst_b200  # noqa: F401
from dvst_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _gemm(A, W, bias=None, residual=None, out_kind=0, act=0, split=0, K=None):
    L = _lib.lib()
    M, N = A.shape[0], W.shape[0]
    K = K or A.shape[1]
    shape = (M, 2 * N) if out_kind == 2 else (M, N)
    out = torch.empty(shape, device=A.device, dtype=torch.float32 if out_kind == 1 else torch.bfloat16)
    _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(bias), M, N, K, split,
                                _lib.ptr(residual), residual.stride(0) if residual is not None else 0,
                                _lib.ptr(out), shape[1], out_kind, N, act, _lib.cur_stream()))
    return out


@pytest.mark.parametrize("M,N,K", [(1, 256, 64), (127, 256, 128), (128, 768, 768), (333, 2304, 768), (4099, 3072, 768),
                                   (2051, 768, 3072), (64, 65536, 256)])
def test_gemm_epilogues(dev, M, N, K):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    ref = A.float() @ W.float().t() + bias
    assert _rel(_gemm(A, W, bias, out_kind=1), ref) < 1e-5          # fp32 out: only summation order differs
    assert _rel(_gemm(A, W, bias, out_kind=0).float(), ref) < 4e-3  # bf16 rounding of the output
    assert _rel(_gemm(A, W, bias, out_kind=0, act=1).float(), torch.nn.functional.gelu(ref)) < 4e-3
    res = torch.randn(M, N, device=dev)
    x = res.clone()
    L = _lib.lib()
    _lib.check(L.dvst_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), M, N, K, 0, _lib.ptr(x), N,
                                _lib.ptr(x), N, 1, 0, 0, _lib.cur_stream()))
    assert _rel(x, ref + res) < 1e-5                                 # in-place residual


def test_gemm_split_precision(dev):
    torch.manual_seed(1)
    M, N, K = 777, 768, 768
    Af, Wf = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev) * 0.05

    def hilo(t):
        hi = t.bfloat16()
        return torch.cat([hi, (t - hi.float()).bfloat16()], 1).contiguous()
    ref = Af.double() @ Wf.double().t()
    out = _gemm(hilo(Af), hilo(Wf), out_kind=1, split=1, K=K)
    assert _rel(out, ref) < 2e-5
    pair = _gemm(hilo(Af), hilo(Wf), out_kind=2, split=1, K=K)
    assert _rel(pair[:, :N].float() + pair[:, N:].float(), ref) < 2e-5
    assert _rel(_gemm(Af.bfloat16(), Wf.bfloat16(), out_kind=1), ref) > 5e-4  # plain bf16 operands are ~100x worse


def test_gemm_rejects_bad_shapes(dev):
    A = torch.zeros(8, 64, device=dev, dtype=torch.bfloat16)
    W = torch.zeros(100, 64, device=dev, dtype=torch.bfloat16)
    with pytest.raises(_lib.DvstError, match="multiple of 256"):
        _gemm(A, W)


def test_layernorm(dev):
    torch.manual_seed(0)
    L = _lib.lib()
    for M in (1, 7, 1000):
        x = torch.randn(M, 768, device=dev) * 3 + 0.5
        g, b = torch.randn(768, device=dev), torch.randn(768, device=dev)
        ref = torch.nn.functional.layer_norm(x, (768,), g, b, 1e-6)
        o32 = torch.empty(M, 768, device=dev)
        _lib.check(L.dvst_layernorm768(_lib.ptr(x), 768, _lib.ptr(g), _lib.ptr(b), _lib.ptr(o32), 768, 1, 0, M, 1e-6,
                                       _lib.cur_stream()))
        assert (o32 - ref).abs().max().item() < 2e-5
        o = torch.empty(M, 1536, device=dev, dtype=torch.bfloat16)
        _lib.check(L.dvst_layernorm768(_lib.ptr(x), 768, _lib.ptr(g), _lib.ptr(b), _lib.ptr(o), 1536, 2, 768, M, 1e-6,
                                       _lib.cur_stream()))
        assert (o[:, :768].float() + o[:, 768:].float() - ref).abs().max().item() < 1e-4


def _attn_ref(q, k, v):
    S = q.shape[0]
    qh, kh, vh = (t.reshape(t.shape[0], 12, 64).transpose(0, 1).float() for t in (q, k, v))
    p = torch.softmax(qh @ kh.transpose(1, 2) * 0.125, dim=-1)
    return (p @ vh).transpose(0, 1).reshape(S, 768)


@pytest.mark.parametrize("mode,B,T,N", [(0, 2, 16, 9), (0, 1, 3, 5), (0, 1, 30, 4), (0, 1, 60, 2), (0, 1, 1, 3),
                                        (1, 2, 3, 196), (1, 1, 2, 36), (2, 2, 2, 196), (3, 2, 2, 196),
                                        (2, 1, 5, 196), (2, 1, 32, 196), (2, 3, 3, 196), (2, 2, 1, 36), (2, 1, 2, 64),
                                        (2, 5, 1, 196)])
def test_attention_modes(dev, mode, B, T, N):
    torch.manual_seed(mode * 100 + T)
    L = _lib.lib()
    Mt = B * N * T
    ncls = B * T if mode == 3 else B
    qkv = (torch.randn(Mt + ncls, 2304, device=dev) * 1.5).bfloat16()
    out = torch.zeros(Mt + ncls, 768, device=dev, dtype=torch.bfloat16)
    cls_out = torch.zeros(B, T, 768, device=dev)
    _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), mode, B, T, N, Mt,
                                _lib.cur_stream()))
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    refo = torch.zeros(Mt + ncls, 768, device=dev)
    refc = torch.zeros(B, T, 768, device=dev)
    rows = torch.arange(Mt, device=dev).reshape(B, N, T)
    for bb in range(B):
        if mode == 0:
            for n in range(N):
                r = rows[bb, n]
                refo[r] = _attn_ref(q[r], k[r], v[r])
        elif mode in (1, 3):
            for t in range(T):
                c = Mt + (bb * T + t if mode == 3 else bb)
                r = torch.cat([torch.tensor([c], device=dev), rows[bb, :, t]])
                o = _attn_ref(q[r], k[r], v[r])
                refo[r[1:]] = o[1:]
                if mode == 1:
                    refc[bb, t] = o[0]
                else:
                    refo[c] = o[0]
        else:
            r = torch.cat([torch.tensor([Mt + bb], device=dev), rows[bb].reshape(-1)])
            refo[r] = _attn_ref(q[r], k[r], v[r])
    if mode == 1:
        assert _rel(cls_out, refc) < 5e-3
        assert _rel(out[:Mt].float(), refo[:Mt]) < 5e-3
        assert float(out[Mt:].float().abs().sum()) == 0  # shared CLS rows are not written by the spatial kernel
    else:
        assert _rel(out.float(), refo) < 5e-3


def test_pair_loss_and_indexed(dev):
    torch.manual_seed(3)
    L = _lib.lib()
    for dim in (768, 65536):
        s, t = torch.randn(9, dim, device=dev), torch.randn(5, dim, device=dev)
        c = torch.randn(dim, device=dev) * 0.1
        si = torch.tensor([0, 1, 8, 3, 3, 2], dtype=torch.int32, device=dev)
        ti = torch.tensor([4, 0, 0, 1, 2, 2], dtype=torch.int32, device=dev)
        out = torch.empty(6, device=dev)
        _lib.check(L.dvst_pair_loss_indexed(_lib.ptr(s), dim, _lib.ptr(si), _lib.ptr(t), dim, _lib.ptr(ti), _lib.ptr(c),
                                            0.04, 0.1, dim, 6, _lib.ptr(out), _lib.cur_stream()))
        ss, tt = s[si.long()].double(), t[ti.long()].double()
        ref = -(torch.softmax((tt - c.double()) / 0.04, -1) * torch.log_softmax(ss / 0.1, -1)).sum(-1)
        assert ((out.double() - ref).abs() / ref.abs()).max().item() < 1e-5


@pytest.mark.parametrize("P,out_dim", [(1, 4096), (37, 65536), (128, 8192), (200, 4096)])
def test_fused_head_loss(dev, P, out_dim):
    """Fused last-layer GEMM + softmax + CE against the materialised fp64 statement on the same bf16 operands."""
    torch.manual_seed(P)
    L = _lib.lib()
    zs = torch.nn.functional.normalize(torch.randn(P, 256, device=dev), dim=-1).bfloat16()
    zt = torch.nn.functional.normalize(torch.randn(P, 256, device=dev), dim=-1).bfloat16()
    w = torch.nn.functional.normalize(torch.randn(out_dim, 256, device=dev), dim=-1).bfloat16()
    c = torch.randn(out_dim, device=dev) * 0.05
    for (tt, ts, cc) in [(0.04, 0.1, c), (0.02, 0.3, None)]:
        loss = torch.empty(P, device=dev)
        nb = L.dvst_head_loss_scratch_bytes(P)
        scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
        _lib.check(L.dvst_head_loss(_lib.ptr(zs), _lib.ptr(zt), 256, _lib.ptr(w), _lib.ptr(cc), tt, ts, P, out_dim,
                                    _lib.ptr(loss), _lib.ptr(scratch), nb, _lib.cur_stream()))
        ls, lt = zs.double() @ w.double().t(), zt.double() @ w.double().t()
        cd = torch.zeros(out_dim, device=dev, dtype=torch.float64) if cc is None else cc.double()
        ref = -(torch.softmax((lt - cd) / tt, -1) * torch.log_softmax(ls / ts, -1)).sum(-1)
        assert ((loss.double() - ref).abs() / ref.abs()).max().item() < 2e-5


def test_colsum_and_center_ema(dev):
    torch.manual_seed(0)
    L = _lib.lib()
    t = torch.randn(6, 4096, device=dev)
    bs = torch.empty(4096, device=dev)
    _lib.check(L.dvst_colsum(_lib.ptr(t), 4096, _lib.ptr(bs), 6, 4096, _lib.cur_stream()))
    assert torch.allclose(bs, t.sum(0), atol=1e-5)
    center = torch.randn(4096, device=dev)
    ref = center * 0.9 + bs / 12.0 * 0.1
    _lib.check(L.dvst_center_ema(_lib.ptr(center), _lib.ptr(bs), 12.0, 0.9, 4096, _lib.cur_stream()))
    assert torch.allclose(center, ref, atol=1e-6)


def test_batched_selection_matches_reference_goldens(dev, golden_dir):
    """dvst_select_frames (one device thread per video, numpy-compatible fp64) vs the index lists the reference's
    FrameSelectionLoader produced (tests/golden/selection.json, 144 videos x N in {8, 16}) and vs the host function on
    long / adversarial curves (ties, flat curves, collisions running off the end of the video)."""
    import json
    import os

    import numpy as np
    from dvst_b200.frame_selection import select_frames_adaptive, select_frames_adaptive_batch
    S = json.load(open(os.path.join(golden_dir, "selection.json")))
    names = sorted(S)
    for N in (8, 16):
        got = select_frames_adaptive_batch([S[n]["loss"] for n in names], [S[n]["n_unsampled"] for n in names], 4, N,
                                           device=dev).cpu().tolist()
        assert got == [S[n][f"idx{N}"] for n in names]
    rng = np.random.default_rng(5)
    curves, nus = [], []
    for F in (3, 7, 8, 9, 127, 128, 129, 450, 1111):
        curves.append((rng.random(F).astype(np.float32) + 0.05).astype(np.float64).tolist()); nus.append(4 * F - 1)
        curves.append([1.0] * F); nus.append(4 * F)                       # flat: exact ties in |cdf - q|
        spike = [1e-6] * F; spike[F // 2] = 5.0                           # all quantiles collide on one frame
        curves.append(spike); nus.append(4 * F - 3)
    for N in (4, 8, 16, 64):
        got = select_frames_adaptive_batch(curves, nus, 4, N, device=dev).cpu().tolist()
        assert got == [select_frames_adaptive(c, nu, 4, N) for c, nu in zip(curves, nus)]


def test_fused_uint8_preprocess_is_bit_exact(dev):
    """dvst_preprocess_u8 vs the oracle's CPU statement of the loader front half (dino_loss_loader.py:80-92): identical
    fp32 operations, so every value must match bit for bit (odd crop offsets and a non-multiple frame count included)."""
    import dvst_oracle as O
    from dvst_b200.dino_similarity import preprocess_frames
    g = torch.Generator().manual_seed(0)
    for (n, H, W, rate, crop) in [(10, 240, 320, 4, 224), (7, 225, 227, 2, 224), (5, 112, 96, 1, 96)]:
        raw = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, generator=g)
        ref = O.preprocess_frames(raw, rate, crop_size=crop)
        got = preprocess_frames(raw.to(dev), rate, crop_size=crop)
        assert got.is_cuda and tuple(got.shape) == tuple(ref.shape)
        assert torch.equal(got.cpu(), ref)


# ---- weighted k-NN evaluation (eval_knn.py:138-178) ----------------------------------------------------------------
def test_knn_similarity_topk_vote_vs_oracle(dev):
    import json
    import os
    import dvst_oracle as O
    from dvst_b200.eval_knn import knn_classifier, knn_classifier_multi, knn_topk, knn_vote
    from dvst_b200.testing import synth_knn_features
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "knn.json")))
    for c in cases:
        trf, trl, tef, tel = synth_knn_features(c["seed"], c["n_train"], c["n_test"], c["dim"], c["num_classes"],
                                                c["noise"])
        top1, top5, rank, idx = O.knn_classifier(trf, trl, tef, tel, c["k"], c["T"], c["num_classes"])
        d, i = knn_topk(trf.to(dev), tef.to(dev), c["k"], rows_per_chunk=128)  # several similarity slabs
        sim = tef @ trf.t()
        # neighbour lists identical except where two fp32 similarities are closer than summation-order noise
        il = i.cpu().long()
        diff = il != idx
        assert diff.float().mean() < 0.01
        assert (torch.gather(sim, 1, il) - torch.gather(sim, 1, idx)).abs()[diff].sum() <= 2e-6 * max(int(diff.sum()), 1)
        assert (d.cpu() - torch.gather(sim, 1, i.cpu().long())).abs().max() < 2e-6
        assert (d[:, :-1] >= d[:, 1:]).all()
        r, pred = knn_vote(d, i, trl.to(dev), tel.to(dev), c["k"], c["T"], c["num_classes"], return_predictions=True)
        assert (r.cpu().long() == rank).float().mean() > 0.99
        got = knn_classifier(trf.to(dev), trl.to(dev), tef.to(dev), tel.to(dev), c["k"], c["T"], c["num_classes"])
        step = 100.0 / c["n_test"]
        # the reference's own numbers (golden) within one test sample
        assert abs(got[0] - c["top1"]) <= step + 1e-9 and abs(got[1] - c["top5"]) <= step + 1e-9, (c, got)
    # one top-k pass serves every k (prefix property) and equals the per-k calls
    trf, trl, tef, tel = [t.to(dev) for t in synth_knn_features(2, 1500, 300, 768, 101, 6.0)]
    multi = knn_classifier_multi(trf, trl, tef, tel, [10, 20, 100, 200], 0.07, 101)
    for k in (10, 200):
        assert multi[k] == knn_classifier(trf, trl, tef, tel, k, 0.07, 101)


def test_knn_topk_ties_and_edges(dev):
    """Equal similarities are taken in ascending index order (duplicated train rows); k = n_train; k = 1."""
    from dvst_b200.eval_knn import knn_topk
    g = torch.Generator().manual_seed(0)
    base = torch.nn.functional.normalize(torch.randn(40, 32, generator=g), dim=1)
    train = base.repeat(8, 1)                       # every similarity value appears 8 times: indices j, j+40, ...
    test = torch.nn.functional.normalize(torch.randn(17, 32, generator=g), dim=1)
    sim = test @ train.t()
    ref = torch.sort(sim, dim=1, descending=True, stable=True)
    for k in (1, 5, 8, 13, 320):
        d, i = knn_topk(train.to(dev), test.to(dev), k)
        i = i.cpu().long()
        # within a group of 8 duplicates the GPU's fp32 sums are identical, so ascending index order must hold
        assert torch.equal(i % 40, ref.indices[:, :k] % 40)
        grp = i // 40
        same = (i[:, 1:] % 40) == (i[:, :-1] % 40)
        assert (grp[:, 1:][same] > grp[:, :-1][same]).all()
        assert len(set(i[0].tolist())) == k
    with pytest.raises(ValueError):
        knn_topk(train.to(dev), test.to(dev), 321)
    with pytest.raises(_lib.DvstError):
        knn_topk(train, test, 3)


# ---- deferred LayerNorm: residual GEMM that also emits bf16(x) + row statistics, and the GEMM that consumes them -------
def _ln_producer(A, W, bias, x, xoff=None):
    L = _lib.lib()
    M, K = A.shape
    xb = torch.empty(M, 768, device=A.device, dtype=torch.bfloat16)
    stats = torch.zeros(M, 6, 2, device=A.device)
    _lib.check(L.dvst_gemm_bf16_ln(_lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(bias), M, 768, K,
                                   None, None, 0.0, _lib.ptr(x), 768, _lib.ptr(xb), 768, _lib.ptr(stats), _lib.ptr(xoff),
                                   None, 0, 0, _lib.cur_stream()))
    return xb, stats


def _ln_consumer(xb, stats, W, b, gamma, beta, act=0):
    L = _lib.lib()
    M, N = xb.shape[0], W.shape[0]
    wf = (W.double() * gamma.double()[None, :]).float().bfloat16().contiguous()
    sf = wf.double().sum(1).float().contiguous()
    bf = (b.double() + W.double() @ beta.double()).float().contiguous()
    out = torch.empty(M, N, device=xb.device, dtype=torch.bfloat16)
    _lib.check(L.dvst_gemm_bf16_ln(_lib.ptr(xb), 768, _lib.ptr(wf), 768, _lib.ptr(bf), M, N, 768, _lib.ptr(stats),
                                   _lib.ptr(sf), 1e-6, None, 0, None, 0, None, None, _lib.ptr(out), N, act,
                                   _lib.cur_stream()))
    return out


@pytest.mark.parametrize("M,K", [(1, 768), (77, 768), (300, 3072), (4099, 768), (40000, 768), (33000, 3072)])
def test_deferred_layernorm_producer(dev, M, K):
    torch.manual_seed(M + K)
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(768, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(768, device=dev)
    x0 = torch.randn(M, 768, device=dev) * 2 + 0.3
    ref = x0.double() + A.double() @ W.double().t() + bias.double()
    for xoff in (None, torch.randn(M, device=dev) * 0.5 + 0.3):
        x = x0.clone()
        xb, stats = _ln_producer(A, W, bias, x, xoff)
        assert _rel(x, ref) < 1e-5                      # in-place fp32 residual stream
        off = 0.0 if xoff is None else xoff[:, None]
        assert torch.equal(xb, (x - off).bfloat16())    # the bf16 copy is the rounding of exactly what was stored, centred
        blocks = (x - off).double().view(M, 6, 128)
        assert torch.allclose(stats[:, :, 0].double(), blocks.sum(-1), rtol=1e-5, atol=1e-3)
        assert torch.allclose(stats[:, :, 1].double(), (blocks * blocks).sum(-1), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("M,N,act", [(1, 2304, 0), (130, 3072, 1), (4099, 2304, 0), (40000, 3072, 1), (40000, 1536, 0)])
def test_deferred_layernorm_consumer(dev, M, N, act):
    torch.manual_seed(M + N)
    L = _lib.lib()
    x = torch.randn(M, 768, device=dev) * 1.7 + 0.2
    x[:, 5] += 9.0                                   # an outlier channel, as real residual streams have
    gamma = torch.randn(768, device=dev) * 0.1 + 1.0
    beta = torch.randn(768, device=dev) * 0.1
    W = torch.randn(N, 768, device=dev) * 0.04
    b = torch.randn(N, device=dev) * 0.1
    ref = torch.nn.functional.layer_norm(x.double(), (768,), gamma.double(), beta.double(), 1e-6) @ W.double().t() + b.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    for centred in (False, True):
        xb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
        stats = torch.empty(M, 6, 2, device=dev)
        xoff = torch.empty(M, device=dev) if centred else None
        _lib.check(L.dvst_rowstats768(_lib.ptr(x), 768, _lib.ptr(xb), 768, _lib.ptr(stats), _lib.ptr(xoff), M,
                                      _lib.cur_stream()))
        c = x - (x.mean(1, keepdim=True) if centred else 0.0)
        if centred:
            assert torch.allclose(xoff, x.mean(1), rtol=1e-5, atol=1e-5)
        assert (xb.float() - c).abs().max() <= 2.0 ** -8 * c.abs().max()
        assert torch.allclose(stats[:, :, 0].double(), c.double().view(M, 6, 128).sum(-1), rtol=1e-4, atol=2e-2)
        out = _ln_consumer(xb, stats, W, b, gamma, beta, act)
        # same error budget as LayerNorm -> bf16 -> GEMM -> bf16: operand and output rounding
        assert _rel(out.float(), ref) < 6e-3


def test_deferred_layernorm_row_offsets_keep_large_mean_rows_accurate(dev):
    """Rows whose mean is 50x their spread: with zero offsets the bf16 copy (rounding relative to |x|) and the
    E[x^2] - mean^2 variance lose the signal; with the offsets moved to the row means (dvst_row_offsets768 between two
    producers) the producer / consumer pair is as accurate as LayerNorm -> bf16 -> GEMM."""
    torch.manual_seed(3)
    L = _lib.lib()
    M, K, N = 4099, 768, 2304
    A = torch.randn(M, K, device=dev).bfloat16()
    Wp = (torch.randn(768, K, device=dev) * 0.03).bfloat16()
    bias = torch.randn(768, device=dev) * 0.1
    x0 = torch.randn(M, 768, device=dev) + 50.0 * torch.randn(M, 1, device=dev).sign()
    gamma = torch.randn(768, device=dev) * 0.1 + 1.0
    beta = torch.randn(768, device=dev) * 0.1
    W = torch.randn(N, 768, device=dev) * 0.04
    b = torch.randn(N, device=dev) * 0.1
    xnew = x0.double() + A.double() @ Wp.double().t() + bias.double()
    ref = torch.nn.functional.layer_norm(xnew, (768,), gamma.double(), beta.double(), 1e-6) @ W.double().t() + b.double()
    # zero offsets: inaccurate (documents why the offsets exist)
    x = x0.clone()
    xb, stats = _ln_producer(A, Wp, bias, x, None)
    bad = _rel(_ln_consumer(xb, stats, W, b, gamma, beta).float(), ref)
    # offsets = means of the previous version of the rows (what the forward does once per block)
    xoff = torch.zeros(M, device=dev)
    xb0 = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    st0 = torch.empty(M, 6, 2, device=dev)
    _lib.check(L.dvst_rowstats768(_lib.ptr(x0), 768, _lib.ptr(xb0), 768, _lib.ptr(st0), None, M, _lib.cur_stream()))
    _lib.check(L.dvst_row_offsets768(_lib.ptr(st0), _lib.ptr(xoff), M, _lib.cur_stream()))
    assert torch.allclose(xoff, x0.mean(1), rtol=1e-4, atol=1e-3)
    x = x0.clone()
    xb, stats = _ln_producer(A, Wp, bias, x, xoff)
    good = _rel(_ln_consumer(xb, stats, W, b, gamma, beta).float(), ref)
    assert good < 6e-3 and bad > 5 * good, (good, bad)


@pytest.mark.parametrize("mode,B,T,N,adversarial", [(1, 5, 5, 196, False), (3, 5, 5, 196, False), (1, 5, 5, 36, False),
                                                     (1, 5, 5, 196, True), (3, 5, 5, 196, True), (1, 5, 5, 36, True),
                                                     (1, 5, 5, 196, "sink"), (3, 5, 5, 196, "sink"), (1, 1, 2, 196, False)])
@pytest.mark.parametrize("impl", [1, 3])
def test_attention_one_pass_softmax_and_fixup(dev, mode, B, T, N, adversarial, impl):
    """tcgen05 spatial attention with scratch: softmax referenced to the maximum of a 32-key sample and the CLS key (the
    split kernel, impl 1: the first 32 keys of its second key half; the first-generation kernel, impl 3: the first 32
    keys).  Ordinary inputs need no fix-up; adversarial ones (row maxima hundreds of nats above the sample: exp2
    overflows) are recomputed exactly by the fix-up kernel - both must match the fp32 reference."""
    torch.manual_seed(7 + mode + N)
    L = _lib.lib()
    npad = (N + 1 + 15) & ~15
    k0 = ((npad // 2) + 15) & ~15
    split = impl == 1 and 32 <= npad <= 208
    # key positions outside the kernel's reference sample
    far_pos = [p for p in range(N) if (p >= 8 and not (k0 <= p < k0 + 32))] if split else list(range(min(40, N - 1), N))
    Mt = B * N * T
    ncls = B * T if mode == 3 else B
    qkv = (torch.randn(Mt + ncls, 2304, device=dev) * 1.5).bfloat16()
    if adversarial == "sink":
        # every query sees the CLS key ~75 nats above all other keys (an attention-sink head): the CLS key is part of the
        # softmax reference, so no row needs the fix-up
        u = torch.nn.functional.normalize(torch.randn(64, device=dev), dim=0)
        for h in range(12):
            qkv[:, h * 64: (h + 1) * 64] += (20.0 * u).bfloat16()
            qkv[Mt:, 768 + h * 64: 768 + (h + 1) * 64] = (30.0 * u).bfloat16()
    elif adversarial:
        # keys whose position inside the frame is >= 40 (and the CLS keys) point along u with a large norm; every third
        # query does too -> scores ~ +-2000 / 8 far above anything among the first 32 keys
        u = torch.nn.functional.normalize(torch.randn(64, device=dev), dim=0)
        rows = torch.arange(Mt, device=dev).reshape(B, N, T)
        far_keys = rows[:, far_pos, :].reshape(-1)
        for h in range(12):
            qkv[far_keys, 768 + h * 64: 768 + (h + 1) * 64] = (45.0 * u).bfloat16()
            qkv[Mt:, 768 + h * 64: 768 + (h + 1) * 64] = (-45.0 * u).bfloat16()
            qkv[0::3, h * 64: (h + 1) * 64] = (45.0 * u).bfloat16()
            qkv[1::7, h * 64: (h + 1) * 64] = (-45.0 * u).bfloat16()
    nb = L.dvst_attention_scratch_bytes(B, T, N)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)
    out = torch.zeros(Mt + ncls, 768, device=dev, dtype=torch.bfloat16)
    cls_out = torch.zeros(B, T, 768, device=dev)
    try:
        L.dvst_set_attention_impl(impl)
        _lib.check(L.dvst_attention_ws(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), mode, B, T, N, Mt,
                                       _lib.ptr(scratch), nb, _lib.cur_stream()))
    finally:
        L.dvst_set_attention_impl(1)
    n_fix = int(scratch[:4].view(torch.int32)[0])
    assert (n_fix > 0) == (adversarial is True)
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    refo = torch.zeros(Mt + ncls, 768, device=dev)
    refc = torch.zeros(B, T, 768, device=dev)
    rows = torch.arange(Mt, device=dev).reshape(B, N, T)
    for bb in range(B):
        for t in range(T):
            c = Mt + (bb * T + t if mode == 3 else bb)
            r = torch.cat([torch.tensor([c], device=dev), rows[bb, :, t]])
            o = _attn_ref(q[r], k[r], v[r])
            refo[r[1:]] = o[1:]
            if mode == 1:
                refc[bb, t] = o[0]
            else:
                refo[c] = o[0]
    assert torch.isfinite(out.float()).all() and torch.isfinite(cls_out).all()
    assert _rel(out[:Mt].float(), refo[:Mt]) < 5e-3
    if mode == 1:
        assert _rel(cls_out, refc) < 5e-3
    else:
        assert _rel(out[Mt:].float(), refo[Mt:]) < 5e-3
    # the two-pass kernel (no scratch) agrees to bf16 rounding of P
    out2 = torch.zeros_like(out)
    cls2 = torch.zeros_like(cls_out)
    _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(out2), 768, _lib.ptr(cls2), mode, B, T, N, Mt,
                                _lib.cur_stream()))
    assert _rel(out.float(), out2.float()) < 5e-3


def test_joint_attention_rescale_path(dev):
    """Joint (flash-style tcgen05) attention with scores that keep growing along the keys, by far more than the lazy
    reference's 2^8 slack per key tile: the O / l rescale runs on (almost) every key tile and must stay exact; a second
    half of the queries sees decreasing scores (no rescale after the first tile)."""
    torch.manual_seed(11)
    L = _lib.lib()
    B, T, N = 2, 4, 196
    NT = N * T
    Mt = B * NT
    qkv = (torch.randn(Mt + B, 2304, device=dev) * 1.0).bfloat16()
    u = torch.nn.functional.normalize(torch.randn(64, device=dev), dim=0)
    tok = torch.arange(NT, device=dev).repeat(B).float()
    ramp = (tok / 128.0).floor()                               # key tile index
    for h in range(12):
        qkv[:Mt, 768 + h * 64: 768 + (h + 1) * 64] += (ramp[:, None] * 9.0 * u[None, :]).bfloat16()
        qkv[0:Mt:2, h * 64: (h + 1) * 64] += (10.0 * u).bfloat16()      # even queries: scores grow ~90 per key tile
        qkv[1:Mt:2, h * 64: (h + 1) * 64] -= (10.0 * u).bfloat16()      # odd queries: scores fall
    out = torch.zeros(Mt + B, 768, device=dev, dtype=torch.bfloat16)
    cls_out = torch.zeros(B, T, 768, device=dev)
    _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(out), 768, _lib.ptr(cls_out), 2, B, T, N, Mt,
                                _lib.cur_stream()))
    q, k, v = qkv[:, :768], qkv[:, 768:1536], qkv[:, 1536:]
    ref = torch.zeros(Mt + B, 768, device=dev)
    for bb in range(B):
        r = torch.cat([torch.tensor([Mt + bb], device=dev), torch.arange(bb * NT, (bb + 1) * NT, device=dev)])
        ref[r] = _attn_ref(q[r], k[r], v[r])
    assert torch.isfinite(out.float()).all()
    assert _rel(out.float(), ref) < 5e-3


@pytest.mark.parametrize("mode,B,T,N,impl", [(1, 2, 3, 196, 1), (3, 2, 2, 196, 1), (1, 2, 3, 36, 1), (1, 3, 2, 100, 1),
                                              (1, 2, 2, 196, 0), (1, 9, 16, 196, 1)])
def test_attention_high_precision_tcgen05(dev, mode, B, T, N, impl):
    """Split-precision spatial attention (bf16 hi / lo operands, hi*hi + hi*lo + lo*hi on tcgen05, P written as a hi / lo
    pair; csrc/attention_hp_sm100.cu) against fp64 torch on the same operands: ~16 mantissa bits end to end.  impl 0 =
    the mma.sync split kernel it replaces.  (9, 16, 196): 1 728 problems, 11 - 12 per CTA (barrier phases wrap many times)."""
    torch.manual_seed(11 + mode + N)
    L = _lib.lib()
    Mt = B * N * T
    ncls = B * T if mode == 3 else B
    x = torch.randn(Mt + ncls, 2304, device=dev) * 1.5
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    qkv = torch.cat([hi, lo], 1).contiguous()                       # [rows, 4608]
    xe = hi.double() + lo.double()                                  # what the kernel's operands represent
    nb = L.dvst_attention_scratch_bytes(B, T, N)
    scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)
    out = torch.zeros(Mt + ncls, 1536, device=dev, dtype=torch.bfloat16)
    cls_out = torch.zeros(B, T, 768, device=dev)
    try:
        L.dvst_set_attention_impl(impl)
        _lib.check(L.dvst_attention_hp(_lib.ptr(qkv), 4608, _lib.ptr(out), 1536, _lib.ptr(cls_out), mode, B, T, N, Mt,
                                       _lib.ptr(scratch), nb, _lib.cur_stream()))
    finally:
        L.dvst_set_attention_impl(1)
    torch.cuda.synchronize()
    got_all = out[:, :768].double() + out[:, 768:].double()
    rows = torch.arange(Mt, device=dev).reshape(B, N, T)
    worst = 0.0
    for bb in range(min(B, 3)):
        for t in range(T if T <= 3 else 2):
            c = Mt + (bb * T + t if mode == 3 else bb)
            r = torch.cat([rows[bb, :, t], torch.tensor([c], device=dev)])
            for h in (0, 5, 11):
                q, k, v = [xe[r, i * 768 + h * 64: i * 768 + h * 64 + 64] for i in range(3)]
                ref = torch.softmax(q @ k.T * 0.125, -1) @ v
                cls = cls_out[bb, t, h * 64:h * 64 + 64].double() if mode == 1 else got_all[c, h * 64:h * 64 + 64]
                got = torch.cat([got_all[r[:-1], h * 64:h * 64 + 64], cls[None]])
                worst = max(worst, ((got - ref).norm() / ref.norm()).item())
    assert worst < 1e-4, worst


@pytest.mark.parametrize("B,T,N", [(2, 16, 9), (1, 3, 50), (1, 30, 7), (1, 1, 3), (3, 32, 5), (1, 48, 3), (1, 60, 2),
                                   (4, 16, 196)])
def test_temporal_attention_high_precision(dev, B, T, N):
    """Split-precision temporal attention (mode 0, bf16 hi / lo operands, every product hi*hi + hi*lo + lo*hi, P split the same
    way; csrc/attention.cu attention_kernel<1, true>) against fp64 torch on the same operands."""
    torch.manual_seed(70 + T)
    L = _lib.lib()
    Mt = B * N * T
    x = torch.randn(Mt + B, 2304, device=dev) * 1.5
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    qkv = torch.cat([hi, lo], 1).contiguous()
    xe = hi.double() + lo.double()
    out = torch.zeros(Mt + B, 1536, device=dev, dtype=torch.bfloat16)
    _lib.check(L.dvst_attention_hp(_lib.ptr(qkv), 4608, _lib.ptr(out), 1536, None, 0, B, T, N, Mt, None, 0,
                                   _lib.cur_stream()))
    torch.cuda.synchronize()
    outs = [out]
    assert float(outs[0][Mt:].float().abs().sum()) == 0          # CLS rows are not part of the temporal attention
    got = outs[0][:Mt, :768].double() + outs[0][:Mt, 768:].double()
    q, k, v = (xe[:Mt, i * 768:(i + 1) * 768].reshape(B * N, T, 12, 64).transpose(1, 2) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(2, 3) * 0.125, -1) @ v).transpose(1, 2).reshape(Mt, 768)
    assert ((got - ref).norm() / ref.norm()).item() < 1e-4


# ---- temporal half of a divided block in one kernel: deferred-LayerNorm qkv GEMM + temporal attention out of shared memory
@pytest.mark.parametrize("T,n_seq,clip_seqs", [(16, 313, 0), (16, 8, 0), (32, 77, 0), (64, 9, 0), (8, 401, 0), (4, 1000, 0),
                                               (1, 300, 0), (30, 131, 0), (3, 1571, 0), (5, 52, 0), (17, 40, 0), (48, 7, 0),
                                               (2, 64, 0), (64, 33, 0), (48, 21, 0), (40, 9, 0), (30, 196 * 3, 196), (3, 196 * 7, 196), (16, 196 * 3, 196),
                                               (5, 36 * 4, 36), (8, 49 * 5, 49)])
def test_fused_temporal_qkv_attention(dev, T, n_seq, clip_seqs):
    """dvst_gemm_ln_temporal_attention vs the two-kernel form (dvst_gemm_bf16_ln consumer -> dvst_attention mode 0) and vs
    an fp64 torch statement of temporal_norm1 -> qkv -> softmax(q k^T / 8) v (models/timesformer.py:127-131, :68-81).
    Tile tails (M not a multiple of the 2 x R tile), sequences that straddle 16-row query tiles (T = 30, 3, 5, 17) and
    one-sequence tiles (T = 48) are covered."""
    torch.manual_seed(T * 1000 + n_seq)
    L = _lib.lib()
    M = T * n_seq
    clip_rows = clip_seqs * T          # 0: one matrix; else rows are tiled clip by clip (clip_seqs sequences per clip)
    x = torch.randn(M, 768, device=dev) * 1.7 + 0.2
    x[:, 5] += 9.0
    gamma = torch.randn(768, device=dev) * 0.1 + 1.0
    beta = torch.randn(768, device=dev) * 0.1
    W = torch.randn(2304, 768, device=dev) * 0.06
    b = torch.randn(2304, device=dev) * 0.1
    xb = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(M, 6, 2, device=dev)
    xoff = torch.empty(M, device=dev)
    _lib.check(L.dvst_rowstats768(_lib.ptr(x), 768, _lib.ptr(xb), 768, _lib.ptr(stats), _lib.ptr(xoff), M,
                                  _lib.cur_stream()))
    # two kernels: packed qkv in HBM, then the temporal attention kernel
    qkv = _ln_consumer(xb, stats, W, b, gamma, beta, 0)
    two = torch.full((M, 768), float("nan"), device=dev, dtype=torch.bfloat16)
    _lib.check(L.dvst_attention(_lib.ptr(qkv), 2304, _lib.ptr(two), 768, None, 0, n_seq, T, 1, M, _lib.cur_stream()))
    # one kernel, head-permuted folded tensors
    wf = (W.double() * gamma.double()[None, :]).float().bfloat16().contiguous()
    sf = wf.double().sum(1).float().contiguous()
    bf = (b.double() + W.double() @ beta.double()).float().contiguous()
    hp = torch.arange(2304, device=dev).view(3, 12, 64).permute(1, 0, 2).reshape(-1)
    wfp, sfp, bfp = wf[hp].contiguous(), sf[hp].contiguous(), bf[hp].contiguous()
    guard = 64
    buf = torch.full((M + guard, 768), 7.0, device=dev, dtype=torch.bfloat16)   # rows past M must stay untouched
    _lib.check(L.dvst_gemm_ln_temporal_attention(_lib.ptr(xb), 768, _lib.ptr(wfp), _lib.ptr(bfp), _lib.ptr(sfp),
                                                 _lib.ptr(stats), 1e-6, M, T, clip_rows, _lib.ptr(buf), 768,
                                                 _lib.cur_stream()))
    torch.cuda.synchronize()
    one = buf[:M]
    assert bool((buf[M:] == 7.0).all())
    assert bool(torch.isfinite(one.float()).all())
    y = torch.nn.functional.layer_norm(x.double(), (768,), gamma.double(), beta.double(), 1e-6) @ W.double().t() + b.double()
    q, k, v = (y[:, i * 768:(i + 1) * 768].view(n_seq, T, 12, 64).transpose(1, 2) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(M, 768)
    assert _rel(one.float(), ref) < 8e-3
    assert _rel(two.float(), ref) < 8e-3
    if T in (16, 32, 64):
        assert torch.equal(one, two)          # same arithmetic in the same order: same bits
    else:
        assert _rel(one.float(), two.float()) < 4e-3
    if clip_rows:
        # clip-by-clip tiling: the output of a clip does not depend on where it sits in the batch
        n_clips = M // clip_rows
        c = n_clips - 1
        sl = slice(c * clip_rows, (c + 1) * clip_rows)
        xb1, st1 = xb[sl].contiguous(), stats[sl].contiguous()
        alone = torch.empty(clip_rows, 768, device=dev, dtype=torch.bfloat16)
        _lib.check(L.dvst_gemm_ln_temporal_attention(_lib.ptr(xb1), 768, _lib.ptr(wfp), _lib.ptr(bfp), _lib.ptr(sfp),
                                                     _lib.ptr(st1), 1e-6, clip_rows, T, clip_rows, _lib.ptr(alone), 768,
                                                     _lib.cur_stream()))
        assert torch.equal(alone, one[sl])


def test_fused_temporal_qkv_attention_rejects_long_sequences(dev):
    L = _lib.lib()
    z = torch.zeros(65 * 2, 768, device=dev, dtype=torch.bfloat16)
    f = torch.zeros(2304, device=dev)
    rc = L.dvst_gemm_ln_temporal_attention(_lib.ptr(z), 768, _lib.ptr(torch.zeros(2304, 768, device=dev).bfloat16()),
                                           _lib.ptr(f), _lib.ptr(f), _lib.ptr(torch.zeros(130, 12, device=dev)), 1e-6,
                                           130, 65, 0, _lib.ptr(z.clone()), 768, _lib.cur_stream())
    assert rc != 0 and b"not covered" in L.dvst_last_error()
