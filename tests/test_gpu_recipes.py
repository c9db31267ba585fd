"""GPU: the recipe layer (preserve_neighbors / preserve_distances / laplacian_embedding, device k-NN,
edge sampling) drives the CUDA path end to end."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _blobs(n, d, k, seed):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((k, d)) * 6
    lab = rng.integers(0, k, n)
    return (centers[lab] + rng.standard_normal((n, d))).astype(np.float32), lab


def test_device_knn_matches_bruteforce():
    import pymde_b200 as pm
    from pymde_b200 import preprocess
    X, _ = _blobs(500, 10, 4, 0)
    g = preprocess.k_nearest_neighbors(X, k=5)
    D = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(D, np.inf)
    nb = np.argsort(D, 1)[:, :5]
    ref = set()
    for i in range(500):
        for j in nb[i]:
            ref.add((min(i, j), max(i, j)))
    got = set(map(tuple, g.edges.tolist()))
    assert len(got ^ ref) <= 2  # ties may differ
    assert set(g.weights.unique().tolist()) <= {1.0, 2.0}  # pymde/test_recipes.py:12-26


def test_preserve_neighbors_end_to_end():
    import pymde_b200 as pm
    X, lab = _blobs(3000, 20, 5, 1)
    pm.seed(0)
    mde = pm.preserve_neighbors(X, embedding_dim=2, verbose=False)
    assert isinstance(mde.distortion_function, pm.penalties.PushAndPull)
    assert mde.constraint is pm.Centered()
    p_att = int((mde.distortion_function.weights > 0).sum())
    assert int(mde.p) <= 2 * p_att and int(mde.p) > 1.8 * p_att  # repulsive fraction 1
    E = mde.embed(max_iter=100)
    st = mde.solve_stats
    assert st.average_distortions[-1] < st.average_distortions[0]
    assert abs(float(E.mean())) < 1e-4
    # neighbourhoods are preserved: same-blob points end up closer than different-blob points
    E = E.cpu().numpy()
    same = np.linalg.norm(E[lab == 0] - E[lab == 0].mean(0), axis=1).mean()
    other = np.linalg.norm(E[lab == 0].mean(0) - E[lab == 1].mean(0))
    assert other > same
    # same seed => identical problem (pymde/test_recipes.py:115-159)
    pm.seed(0)
    mde2 = pm.preserve_neighbors(X, embedding_dim=2)
    assert torch.equal(mde.edges, mde2.edges)
    assert torch.equal(mde.distortion_function.weights, mde2.distortion_function.weights)


def test_laplacian_embedding_standardized():
    import pymde_b200 as pm
    X, _ = _blobs(800, 8, 3, 2)
    mde = pm.laplacian_embedding(X, embedding_dim=2)
    assert mde.constraint is pm.Standardized()
    E = mde.embed(max_iter=60)
    E64 = E.double()
    np.testing.assert_allclose((E64.T @ E64 / 800).cpu().numpy(), np.eye(2), atol=1e-4)


def test_preserve_distances_cycle_graph_c1():
    """BASELINE config C1: 1k-node cycle graph, preserve_distances, d=2 (Absolute loss, Centered)."""
    import pymde_b200 as pm
    n = 1000
    g = pm.preprocess.Graph.from_edges(torch.tensor([(i, (i + 1) % n) for i in range(n)]))
    pm.seed(0)
    mde = pm.preserve_distances(g, embedding_dim=2)
    assert int(mde.p) == n * (n - 1) // 2
    assert isinstance(mde.distortion_function, pm.losses.Absolute)
    X0 = mde.constraint.initialization(n, 2, mde.device)
    v0 = mde.average_distortion(X0).item()
    E = mde.embed(X=X0, max_iter=150)
    v = mde.average_distortion(E).item()
    assert v < 0.25 * v0
    # a cycle embeds as (roughly) a circle: radii concentrate
    r = E.norm(dim=1)
    assert float(r.std() / r.mean()) < 0.2


def test_preserve_distances_standardized_scales_deviations():
    import pymde_b200 as pm
    X, _ = _blobs(300, 6, 3, 3)
    mde = pm.preserve_distances(X, embedding_dim=2, constraint=pm.Standardized(), loss=pm.losses.Quadratic)
    d = mde.distortion_function.deviations
    nat = float(pm.Standardized().natural_length(300, 2))
    np.testing.assert_allclose(float(d.pow(2).mean().sqrt()), nat, rtol=1e-5)
    mde.embed(max_iter=30)
    assert mde.solve_stats.average_distortions[-1] < mde.solve_stats.average_distortions[0]
