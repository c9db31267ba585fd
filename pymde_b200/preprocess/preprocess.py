"""Edge sampling and scaling utilities (interface of pymde/preprocess/preprocess.py)."""
import numpy as np
import torch

from .. import util


def _keys(e, n):
    lo = torch.minimum(e[:, 0], e[:, 1])
    hi = torch.maximum(e[:, 0], e[:, 1])
    return lo * n + hi


def sample_edges(n, num_edges, exclude=None, seed=None, device=None):
    """Uniformly sample (at most) `num_edges` distinct pairs i < j, none of them in `exclude`.

    Pairs are drawn on the device, canonicalised, de-duplicated in draw order and filtered against the
    excluded set with a sorted search on 64-bit keys (the reference does the same with a
    triangular-number bijection + np.unique on the host, preprocess.py:11-80).  Like the reference the
    result may hold fewer than `num_edges` rows."""
    n = int(n)
    num_edges = int(num_edges)
    n_all = n * (n - 1) // 2
    n_excl = 0 if exclude is None else int(exclude.shape[0])
    if num_edges > n_all - n_excl:
        raise ValueError("Cannot sample more than (%d choose 2) - %d = %d edges. (requested: %d edges)"
                         % (n, n_excl, n_all - n_excl, num_edges))
    dev = torch.device("cpu") if device is None and not torch.cuda.is_available() else util.cuda_device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) if seed is not None else int(util.np_rng().integers(0, 2 ** 62)))
    excl = None
    if exclude is not None and n_excl:
        ex = exclude if isinstance(exclude, torch.Tensor) else torch.as_tensor(np.asarray(exclude))
        excl = torch.sort(_keys(ex.to(dev).long(), n)).values
    got = torch.empty(0, dtype=torch.int64, device=dev)
    draws = 0
    while got.numel() < num_edges and draws < 64:
        want = num_edges - got.numel()
        m = int(want * 1.15) + 1024
        e = torch.randint(0, n, (m, 2), generator=gen, device=dev)
        e = e[e[:, 0] != e[:, 1]]
        key = _keys(e, n)
        if excl is not None:
            pos = torch.searchsorted(excl, key).clamp_(max=excl.numel() - 1)
            key = key[excl[pos] != key]
        key = torch.cat([got, key])
        # de-duplicate keeping first occurrences (stable draw order)
        srt, order = torch.sort(key, stable=True)
        first = torch.ones_like(srt, dtype=torch.bool)
        first[1:] = srt[1:] != srt[:-1]
        keep = torch.sort(order[first]).values
        got = key[keep][:num_edges]
        draws += 1
    return torch.stack([got // n, got % n], 1)


def dissimilar_edges(n_items, similar_edges, num_edges=None, seed=None):
    """Edges NOT in `similar_edges`, approximately as many as there are similar ones."""
    if num_edges is None:
        num_edges = similar_edges.shape[0]
    return sample_edges(n_items, num_edges, exclude=similar_edges, seed=seed)


def deduplicate_edges(edges):
    e = edges if isinstance(edges, torch.Tensor) else torch.as_tensor(np.asarray(edges))
    lo, hi = torch.minimum(e[:, 0], e[:, 1]), torch.maximum(e[:, 0], e[:, 1])
    return torch.unique(torch.stack([lo, hi], 1), dim=0)


def scale(distances, natural_length):
    """Rescale so that RMS(distances) == natural_length (preprocess.py:132-138)."""
    rms = distances.float().pow(2).mean().sqrt()
    return (float(natural_length) / rms) * distances
