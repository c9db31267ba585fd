"""Problem construction from a data matrix (interface of pymde/preprocess/data_matrix.py).

k-nearest neighbours and pairwise distances are computed EXACTLY on the GPU in row chunks (one
library GEMM per chunk for the cross terms, then top-k) -- the reference uses scikit-learn brute
force below 10 000 rows and the approximate pynndescent above (data_matrix.py:125-143)."""
import numpy as np
import scipy.sparse as sp
import torch

from .. import util
from .graph import Graph
from .preprocess import sample_edges  # noqa: F401  (the reference exposes it here as well)


def _to_device_matrix(data, device):
    if sp.issparse(data):
        data = data.toarray()
    if isinstance(data, np.ndarray):
        data = torch.from_numpy(np.ascontiguousarray(data))
    return data.to(device=device, dtype=torch.float32)


def k_nearest_neighbors(data, k, max_distance=None, verbose=False, device=None, chunk_rows=None):
    """Graph whose edges join each row to its k nearest rows (Euclidean); reciprocal pairs get weight 2."""
    dev = util.cuda_device(device)
    X = _to_device_matrix(data, dev)
    n = X.shape[0]
    k = int(min(k, n - 1))
    sq = (X * X).sum(1)
    rows = chunk_rows or max(256, min(n, int(2 ** 27 // max(n, 1))))
    src, dst = [], []
    for s0 in range(0, n, rows):
        Q = X[s0:s0 + rows]
        d2 = (sq[s0:s0 + rows, None] + sq[None, :] - 2.0 * (Q @ X.T)).clamp_(min=0)
        d2[torch.arange(Q.shape[0], device=dev), torch.arange(s0, s0 + Q.shape[0], device=dev)] = float("inf")
        val, idx = torch.topk(d2, k, dim=1, largest=False)
        keep = torch.ones_like(val, dtype=torch.bool) if max_distance is None else val.sqrt() <= max_distance
        i = torch.arange(s0, s0 + Q.shape[0], device=dev)[:, None].expand_as(idx)
        src.append(i[keep]); dst.append(idx[keep])
    e = torch.stack([torch.cat(src), torch.cat(dst)], 1).cpu()
    return Graph.from_edges(e, None, n_items=n)


def distances(data, retain_fraction=1.0, verbose=False, device=None):
    """Graph of pairwise Euclidean distances: all (n choose 2) pairs, or a uniform sample of them."""
    dev = util.cuda_device(device)
    X = _to_device_matrix(data, dev)
    n = X.shape[0]
    n_all = n * (n - 1) // 2
    if retain_fraction >= 1.0:
        edges = torch.triu_indices(n, n, 1, device=dev).T
    else:
        edges = sample_edges(n, int(retain_fraction * n_all), device=dev)
    out = torch.empty(edges.shape[0], dtype=torch.float32, device=dev)
    step = 1 << 22
    for s0 in range(0, edges.shape[0], step):
        e = edges[s0:s0 + step]
        out[s0:s0 + step] = (X[e[:, 0]] - X[e[:, 1]]).norm(dim=1)
    g = Graph.from_edges(edges.cpu(), out.cpu(), n_items=n)
    return g
