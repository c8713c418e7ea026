"""Drop-in for the evaluation half of the reference's eval_knn.py: frozen-backbone feature extraction and the weighted
k-NN classifier, both on the device.

  * `extract_features` - eval_knn.py:94-135 (sharded batches -> [len(dataset), 768] on rank 0; see parallel.py).
  * `knn_classifier`   - eval_knn.py:138-178, same signature and return value (top1, top5 in percent).
  * `knn_topk` / `knn_vote` - the two halves, so that the reference's loop over `--nb_knn 10 20 100 200`
                         (eval_knn.py:229-233) can reuse ONE similarity / top-k pass for every k.

Callers L2-normalise the features first exactly like the reference does (eval_knn.py:69-70: nn.functional.normalize).
"""
import torch

from . import _lib
from .parallel import extract_features  # noqa: F401  (re-exported under the reference's module name)


def _f32c(t, name):
    if not t.is_cuda:
        raise _lib.DvstError(f"dvst_b200 runs on CUDA (sm_100a) only - there is no CPU fallback ({name} is on the CPU)")
    return t.detach().to(torch.float32).contiguous()


def knn_topk(train_features, test_features, k, rows_per_chunk=8192):
    """(distances fp32 [n_test, k], indices int32 [n_test, k]) = (test @ train.T).topk(k, largest=True, sorted=True);
    equal similarities are ordered by ascending train index.  The [rows_per_chunk, n_train] similarity slab is the only
    temporary."""
    train = _f32c(train_features, "train_features")
    test = _f32c(test_features, "test_features")
    n_train, dim = train.shape
    n_test = test.shape[0]
    if test.shape[1] != dim:
        raise ValueError("train / test feature dims differ")
    if not 1 <= k <= min(n_train, 512):
        raise ValueError(f"k={k} must be in [1, min(n_train, 512)]")
    if dim % 4:
        raise ValueError("feature dim must be a multiple of 4")
    dev = test.device
    dist = torch.empty(n_test, k, dtype=torch.float32, device=dev)
    idx = torch.empty(n_test, k, dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        rows = max(1, min(rows_per_chunk, n_test))
        sim = torch.empty(rows, n_train, dtype=torch.float32, device=dev)
        st = _lib.cur_stream()
        for r0 in range(0, n_test, rows):
            nr = min(rows, n_test - r0)
            _lib.check(L.dvst_knn_similarity(_lib.ptr(test[r0:]), dim, _lib.ptr(train), dim, nr, n_train, dim,
                                             _lib.ptr(sim), n_train, st))
            _lib.check(L.dvst_knn_topk(_lib.ptr(sim), n_train, nr, n_train, k, _lib.ptr(dist[r0:]), _lib.ptr(idx[r0:]),
                                       k, st))
    return dist, idx


def knn_vote(distances, indices, train_labels, test_labels, k, T, num_classes=1000, return_predictions=False):
    """rank int32 [n_test]: position of the true class in the weighted vote of the first k neighbours
    (probs.sort(descending), ties by ascending class index).  top1 = (rank == 0), top5 = (rank < 5)."""
    if not distances.is_cuda:
        raise _lib.DvstError("dvst_b200 runs on CUDA (sm_100a) only - there is no CPU fallback")
    dev = distances.device
    n_test, ldk = distances.shape
    if indices.shape != distances.shape or indices.dtype != torch.int32 or distances.dtype != torch.float32:
        raise ValueError("distances fp32 / indices int32 of the same [n_test, k_max] shape expected (knn_topk output)")
    if not (distances.is_contiguous() and indices.is_contiguous()):
        raise ValueError("distances / indices must be contiguous")
    trl = train_labels.detach().to(device=dev, dtype=torch.int64).contiguous()
    tel = test_labels.detach().to(device=dev, dtype=torch.int64).contiguous()
    if tel.numel() != n_test:
        raise ValueError("one test label per test feature expected")
    rank = torch.empty(n_test, dtype=torch.int32, device=dev)
    pred = torch.empty(n_test, dtype=torch.int32, device=dev) if return_predictions else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dvst_knn_vote(_lib.ptr(distances), _lib.ptr(indices), ldk, _lib.ptr(trl), _lib.ptr(tel),
                                            n_test, int(k), float(T), int(num_classes), _lib.ptr(rank), _lib.ptr(pred),
                                            _lib.cur_stream()))
    return (rank, pred) if return_predictions else rank


@torch.no_grad()
def knn_classifier(train_features, train_labels, test_features, test_labels, k, T, num_classes=1000):
    """eval_knn.knn_classifier (eval_knn.py:138-178): returns (top1, top5) accuracies in percent."""
    dist, idx = knn_topk(train_features, test_features, k)
    rank = knn_vote(dist, idx, train_labels, test_labels, k, T, num_classes)
    total = rank.numel()
    top1 = float((rank == 0).sum().item())
    top5 = float((rank < 5).sum().item())
    return top1 * 100.0 / total, top5 * 100.0 / total


@torch.no_grad()
def knn_classifier_multi(train_features, train_labels, test_features, test_labels, ks, T, num_classes=1000):
    """{k: (top1, top5)} for every k of `ks` from one similarity / top-k pass (the reference recomputes it per k)."""
    ks = [int(k) for k in ks]
    dist, idx = knn_topk(train_features, test_features, max(ks))
    out = {}
    for k in ks:
        rank = knn_vote(dist, idx, train_labels, test_labels, k, T, num_classes)
        out[k] = (float((rank == 0).sum().item()) * 100.0 / rank.numel(),
                  float((rank < 5).sum().item()) * 100.0 / rank.numel())
    return out
