"""Quadratic (spectral) initialisation (interface of pymde/quadratic.py:16-179): the bottom eigenvectors
of the graph Laplacian, standardised.  Runs once before `embed`; sparse Lanczos on the host (scipy)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg
import torch

from . import util


def pca(Y, embedding_dim):
    """Top principal directions, scaled to a standardized embedding (quadratic.py:16-44)."""
    Y = Y if isinstance(Y, torch.Tensor) else torch.as_tensor(np.asarray(Y))
    Y = Y.float()
    n = Y.shape[0]
    Yc = Y - Y.mean(0)
    U, _, _ = torch.linalg.svd(Yc, full_matrices=False)
    return (n ** 0.5) * U[:, :embedding_dim]


def _laplacian(n, edges, weights):
    e = edges.detach().cpu().numpy() if isinstance(edges, torch.Tensor) else np.asarray(edges)
    w = weights.detach().cpu().numpy() if isinstance(weights, torch.Tensor) else np.asarray(weights)
    A = sp.coo_matrix((w, (e[:, 0], e[:, 1])), shape=(n, n), dtype=np.float64)
    A = (A + A.T).tocsr()
    return sp.diags(np.asarray(A.sum(1)).ravel()) - A


def spectral(n_items, embedding_dim, edges, weights, cg=False, max_iter=1000, device=None):
    """Standardized spectral embedding: eigenvectors 2..m+1 of L = D - W (quadratic.py:122-179)."""
    L = _laplacian(int(n_items), edges, weights)
    k = int(embedding_dim) + 1
    rng = np.random.default_rng(0)
    if int(n_items) <= max(k + 1, 32):  # Lanczos needs k < n; tiny problems go through a dense eigensolver
        vals, vecs = np.linalg.eigh(L.toarray())
        vals, vecs = vals[:k], vecs[:, :k]
        if vecs.shape[1] < k:  # fewer items than requested directions: pad with random columns
            vecs = np.concatenate([vecs, rng.standard_normal((int(n_items), k - vecs.shape[1]))], 1)
            vals = np.concatenate([vals, np.full(k - vals.shape[0], np.inf)])
        return _finish(vals, vecs, k, n_items, device)
    try:
        vals, vecs = scipy.sparse.linalg.eigsh(L, k=k, sigma=-1e-3 * max(1.0, L.diagonal().mean()), which="LM",
                                               maxiter=max_iter, v0=rng.standard_normal(int(n_items)))
    except Exception:
        vals, vecs = scipy.sparse.linalg.eigsh(L, k=k, which="SA", maxiter=max_iter * 10,
                                               v0=rng.standard_normal(int(n_items)))
    return _finish(vals, vecs, k, n_items, device)


def _finish(vals, vecs, k, n_items, device):
    order = np.argsort(vals)
    X = torch.tensor(vecs[:, order[1:k]].astype(np.float32))
    if torch.cuda.is_available():
        X = X.to(util.cuda_device(device)).contiguous()
        return util.proj_standardized(X, demean=True, inplace=True)
    X = X - X.mean(0)
    U, _, Vh = torch.linalg.svd(X, full_matrices=False)
    return (float(n_items) ** 0.5) * (U @ Vh)
