"""Quadratic (spectral) initialisation (interface of pymde/quadratic.py:16-179): the bottom eigenvectors
of the graph Laplacian, standardised.  Runs once before `embed`.

On a CUDA device the eigenvectors come from `spectral_device`: a block LOBPCG iteration whose only n-sized
sparse operation, V -> L V, IS the gradient scatter of the edge kernels (`mde_scatter_external` with the edge
weights as per-edge coefficients: sum_k w_k (v_i - v_j)(e_i - e_j) = L V), so the initialisation reuses the
layout machinery of the hot path and never leaves the device.  The host path (scipy Lanczos, what the
reference does at quadratic.py:84-96) remains for tiny problems and as the parity arbiter of the tests."""
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


class _LaplacianOperator(object):
    """V (n, k) -> L V on the device through the edge layout (k <= 4: quad / tile kernels, else the wide kernel)."""

    def __init__(self, n, k, edges, weights, device):
        from . import _lib
        from .problem import EdgeLayout
        table = _lib.mde_fn_t()
        table.fn_att = table.fn_rep = 100  # MDE_FN_EXTERNAL: the layout only carries the index structure
        self.w = util.as_f32_cuda(weights, device).reshape(-1).contiguous()
        zeros = torch.zeros(int(edges.shape[0]), device=device)
        self.layout = EdgeLayout(edges, int(n), table, zeros, None, device, embedding_dim=int(k))
        e = edges.to(device)
        deg = torch.zeros(int(n), device=device)
        deg.index_add_(0, e[:, 0], self.w)
        deg.index_add_(0, e[:, 1], self.w)
        self.degree = deg

    def __call__(self, V):
        return self.layout.scatter_external(V.contiguous(), self.w)


def lobpcg_smallest(apply_A, n, k, precond=None, device=None, max_iter=300, tol=1e-4, seed=0, deflate_constant=True,
                    a_norm=1.0, n_wanted=None):
    """k smallest eigenpairs of a symmetric positive semi-definite operator by block LOBPCG (Knyazev 2001), fp32
    vectors, fp64 Rayleigh-Ritz.  `deflate_constant`: iterate in the orthogonal complement of the all-ones vector (the
    Laplacian's known null vector).  Returns (eigenvalues (k,), eigenvectors (n, k), iterations, residual norms)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    def clean(V):
        return V - V.mean(0, keepdim=True) if deflate_constant else V

    def ortho(V):
        Q, _ = torch.linalg.qr(V.double())
        return Q.float()

    X = ortho(clean(torch.randn(n, k, device=device, generator=gen)))
    AX = apply_A(X)
    lam, C = torch.linalg.eigh((X.double().T @ AX.double()))
    X, AX = (X.double() @ C).float(), (AX.double() @ C).float()
    P = AP = None
    res = None
    it = 0
    for it in range(1, max_iter + 1):
        R = AX - X * lam.float()[None, :]
        res = R.norm(dim=0)
        # ||A x - lambda x|| <= tol * lambda (eigsh's relative criterion, quadratic.py:91), floored at the fp32
        # resolution of the operator, 2e-6 * ||A||
        nw = k if n_wanted is None else int(n_wanted)
        if bool((res <= torch.maximum(tol * lam.abs().float(), torch.full_like(res, 2e-6 * a_norm)))[:nw].all()):
            break
        W = R if precond is None else R * precond[:, None]
        W = clean(W)
        W = W - X @ (X.T @ W)
        if P is not None:
            W = W - P @ torch.linalg.lstsq(P.double().T @ P.double(), (P.T @ W).double()).solution.float()
        W = ortho(W)
        AW = apply_A(W)
        S = torch.cat([X, W] + ([P] if P is not None else []), 1)
        AS = torch.cat([AX, AW] + ([AP] if AP is not None else []), 1)
        Sd, ASd = S.double(), AS.double()
        B = Sd.T @ Sd
        G = Sd.T @ ASd
        G = 0.5 * (G + G.T)
        # generalized symmetric eigenproblem through the Cholesky factor of the (well conditioned) Gram matrix;
        # when the basis has become numerically dependent drop P and restart the recurrence
        try:
            Lc = torch.linalg.cholesky(B)
        except Exception:
            P = AP = None
            continue
        Gt = torch.linalg.solve_triangular(Lc, torch.linalg.solve_triangular(Lc, G, upper=False).T, upper=False).T
        Gt = 0.5 * (Gt + Gt.T)
        ev, Y = torch.linalg.eigh(Gt)
        Cc = torch.linalg.solve_triangular(Lc.T, Y[:, :k], upper=True)
        lam = ev[:k]
        Cx, Crest = Cc[:k], Cc[k:]
        rest, Arest = Sd[:, k:], ASd[:, k:]
        Pn, APn = rest @ Crest, Arest @ Crest
        X = (Sd[:, :k] @ Cx + Pn).float()
        AX = (ASd[:, :k] @ Cx + APn).float()
        P, AP = Pn.float(), APn.float()
    return lam, X, it, res


def spectral_device(n_items, embedding_dim, edges, weights, device, max_iter=300, tol=1e-4):
    """Device path of `spectral`: eigenvectors 2..m+1 of L = D - W by LOBPCG with the Jacobi (degree) preconditioner,
    the constant vector deflated exactly (quadratic.py:71-120 asks eigsh / torch.lobpcg for m + 1 vectors and drops
    the first)."""
    n, m = int(n_items), int(embedding_dim)
    dev = util.cuda_device(device)
    edges = edges if isinstance(edges, torch.Tensor) else torch.as_tensor(np.asarray(edges))
    kb = min(m + 2, n - 2)  # two guard vectors: 10-100x fewer iterations on poorly separated spectra
    op = _LaplacianOperator(n, kb, edges.to(dev), weights, dev)
    precond = 1.0 / op.degree.clamp_min(1e-12)
    lam, X, iters, res = lobpcg_smallest(op, n, kb, precond=precond, device=dev, max_iter=max_iter, tol=tol,
                                         a_norm=2.0 * float(op.degree.max()), n_wanted=m)
    lam, res = lam[:m], res[:m]
    X = X[:, :m].contiguous()
    out = util.proj_standardized(X, demean=True, inplace=True)
    out._lobpcg_info = {"iterations": iters, "eigenvalues": lam.cpu().numpy(), "residuals": res.cpu().numpy()}
    return out


def spectral(n_items, embedding_dim, edges, weights, cg=False, max_iter=1000, device=None):
    """Standardized spectral embedding: eigenvectors 2..m+1 of L = D - W (quadratic.py:122-179).  CUDA problems with
    more than 2 000 items use the device LOBPCG (`spectral_device`); PYMDE_B200_SPECTRAL=host forces the host path."""
    import os
    if (torch.cuda.is_available() and int(n_items) > 2000 and os.environ.get("PYMDE_B200_SPECTRAL", "device") != "host"
            and int(embedding_dim) < int(n_items) - 2):
        return spectral_device(n_items, embedding_dim, edges, weights, device, max_iter=min(int(max_iter), 400))
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
