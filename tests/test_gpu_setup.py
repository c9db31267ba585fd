"""SURVEY section 8 row f4 on the device: spectral (quadratic) initialisation by LOBPCG with the edge kernel as the
Laplacian operator, and hop-count shortest paths by bit-parallel multi-source BFS.  Arbiters: scipy's Lanczos
(what the reference calls, pymde/quadratic.py:84-96) and scipy.sparse.csgraph (exact hop counts)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.csgraph as csgraph
import scipy.sparse.linalg as sla
import torch

pytestmark = pytest.mark.gpu


def _geometric_knn(n, k, seed, aspect=(3.0, 1.0)):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2)) * np.array(aspect)
    _, idx = cKDTree(pts).query(pts, k=k + 1)
    e = np.stack([np.repeat(np.arange(n), k), idx[:, 1:].ravel()], 1)
    return np.unique(np.sort(e, axis=1), axis=0).astype(np.int64)


@pytest.mark.parametrize("m", [2, 3])
def test_spectral_device_matches_lanczos(m):
    import pymde_b200 as pm
    from pymde_b200 import quadratic
    n = 20000
    e = _geometric_knn(n, 8, 0)
    w = np.ones(len(e), np.float32)
    X = quadratic.spectral_device(n, m, torch.tensor(e), torch.tensor(w), "cuda")
    info = X._lobpcg_info
    L = quadratic._laplacian(n, e, w)
    vals, vecs = sla.eigsh(L, k=m + 1, sigma=-1e-3, which="LM")
    order = np.argsort(vals)
    ref_vals, Q = vals[order][1:], np.linalg.qr(vecs[:, order[1:]])[0]
    np.testing.assert_allclose(info["eigenvalues"], ref_vals, rtol=2e-3)
    Xn = np.linalg.qr(X.double().cpu().numpy())[0]
    sines = np.linalg.svd(Xn - Q @ (Q.T @ Xn), compute_uv=False)  # sines of the principal angles
    assert sines.max() < 2e-2, sines
    # standardized and centred, like quadratic.py:178-179
    Xc = X.double()
    assert float(Xc.mean(0).abs().max()) < 1e-4
    np.testing.assert_allclose((Xc.T @ Xc / n).cpu().numpy(), np.eye(m), atol=1e-3)
    assert info["iterations"] < 400


def test_spectral_dispatch_uses_device_and_host_agrees(monkeypatch):
    import pymde_b200 as pm
    from pymde_b200 import quadratic
    n, m = 6000, 2
    e = _geometric_knn(n, 8, 1)
    w = torch.ones(len(e))
    Xd = quadratic.spectral(n, m, torch.tensor(e), w, device="cuda")
    assert hasattr(Xd, "_lobpcg_info")
    monkeypatch.setenv("PYMDE_B200_SPECTRAL", "host")
    Xh = quadratic.spectral(n, m, torch.tensor(e), w, device="cuda")
    assert not hasattr(Xh, "_lobpcg_info")
    # same subspace: the quadratic objective of both initialisations agrees
    mde = pm.MDE(n, m, torch.tensor(e, device="cuda"), pm.penalties.Quadratic(w.cuda()), pm.Standardized())
    vd, vh = mde.average_distortion(Xd).item(), mde.average_distortion(Xh).item()
    np.testing.assert_allclose(vd, vh, rtol=2e-3)


def test_hops_match_csgraph_exactly():
    from pymde_b200.preprocess import graph as G
    n = 3000
    e = _geometric_knn(n, 4, 2, aspect=(1.0, 1.0))
    A = sp.coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(n, n))
    A = ((A + A.T) > 0).astype(np.float64).tocsr()
    g = G.Graph(A)
    out = G.shortest_paths_device(g, retain_fraction=1.0, device="cuda")
    D = csgraph.shortest_path(A, directed=False, unweighted=True)
    iu = np.triu_indices(n, 1)
    finite = np.isfinite(D[iu])
    want_edges = np.stack([iu[0][finite], iu[1][finite]], 1)
    want_len = D[iu][finite].astype(np.float32)
    got_edges = out.edges.cpu().numpy()
    assert got_edges.shape == want_edges.shape
    assert np.array_equal(got_edges, want_edges)          # sorted by (i, j), every reachable pair once
    assert np.array_equal(out.distances.cpu().numpy(), want_len)   # exact hop counts


def test_hops_sampling_and_limit():
    from pymde_b200.preprocess import graph as G
    n = 5000
    e = _geometric_knn(n, 5, 3, aspect=(1.0, 1.0))
    A = sp.coo_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(n, n))
    A = ((A + A.T) > 0).astype(np.float64).tocsr()
    g = G.Graph(A)
    full = G.shortest_paths_device(g, max_length=6, retain_fraction=1.0, device="cuda", seed=7)
    assert float(full.distances.max()) <= 6
    D = csgraph.dijkstra(A, directed=False, unweighted=True, limit=6)
    iu = np.triu_indices(n, 1)
    assert full.n_edges == int(np.isfinite(D[iu]).sum())
    part = G.shortest_paths_device(g, max_length=6, retain_fraction=0.25, device="cuda", seed=7)
    again = G.shortest_paths_device(g, max_length=6, retain_fraction=0.25, device="cuda", seed=7)
    assert torch.equal(part.edges, again.edges) and torch.equal(part.distances, again.distances)  # same seed, same sample
    frac = part.n_edges / full.n_edges
    assert abs(frac - 0.25) < 0.01
    # the sample is a subset with the same lengths
    key_full = full.edges[:, 0] * n + full.edges[:, 1]
    key_part = part.edges[:, 0] * n + part.edges[:, 1]
    pos = torch.searchsorted(key_full, key_part)
    assert torch.equal(key_full[pos], key_part) and torch.equal(full.distances[pos], part.distances)


def test_preserve_distances_on_a_graph_stays_on_device():
    import pymde_b200 as pm
    from pymde_b200.preprocess import graph as G
    n = 4000
    e = _geometric_knn(n, 4, 4, aspect=(1.0, 1.0))
    g = pm.preprocess.Graph.from_edges(e, n_items=n)
    pm.seed(0)
    mde = pm.preserve_distances(g, embedding_dim=2, max_distances=2e6, device="cuda")
    assert int(mde.p) > 1.5e6 and mde.edges.device.type == "cuda"
    X = mde.embed(max_iter=30)
    st = mde.solve_stats
    assert st.average_distortions[-1] < st.average_distortions[0]
