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


def test_knn_tiny_known_answer():
    # pymde/test_recipes.py:12-26
    from pymde_b200 import preprocess
    g = preprocess.data_matrix.k_nearest_neighbors(np.array([[0.0], [1.0], [1.5], [1.75]]), k=2)
    assert g.edges.tolist() == [[0, 1], [0, 2], [1, 2], [1, 3], [2, 3]]
    assert g.weights.tolist() == [1.0, 1.0, 2.0, 2.0, 2.0]


def test_standardized_initialization():
    # pymde/test_optim.py:13-18
    import pymde_b200 as pm
    torch.manual_seed(0)
    X = pm.Standardized().initialization(5, 3, device="cuda")
    np.testing.assert_allclose((X.T @ X / 5).cpu().numpy(), np.eye(3), atol=1e-5)


def test_docs_example_reaches_6_2884():
    # docs_src/source/mde/index.rst:193-286: 5 items, 4 edges, Quadratic [1,2,5,6], Standardized
    import pymde_b200 as pm
    edges = torch.tensor([[0, 1], [0, 4], [1, 2], [2, 3]])
    weights = torch.tensor([1.0, 2.0, 5.0, 6.0])
    for s in range(3):
        torch.manual_seed(s)
        mde = pm.MDE(n_items=5, embedding_dim=2, edges=edges, distortion_function=pm.penalties.Quadratic(weights),
                     constraint=pm.Standardized())
        E = mde.embed()
        np.testing.assert_allclose(mde.average_distortion(E).item(), 6.2884, rtol=2e-4)
        np.testing.assert_allclose((E.T @ E / 5).cpu().numpy(), np.eye(2), atol=1e-4)
        assert float(E.mean(0).abs().max()) < 1e-5


def test_laplacian_embedding_is_quadratic_preserve_neighbors():
    # pymde/test_recipes.py:30-51: the two constructions agree after a Procrustes alignment
    import pymde_b200 as pm
    torch.manual_seed(0)
    Y = torch.randn(100, 10)
    pm.seed(0)
    a = pm.laplacian_embedding(Y).embed()
    pm.seed(0)
    b = pm.preserve_neighbors(Y, attractive_penalty=pm.penalties.Quadratic, repulsive_penalty=None).embed()
    b = pm.util.align(source=b, target=a)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-3, atol=5e-3)


def test_anchor_initialization_and_no_anchor_anchor_edges():
    # pymde/test_recipes.py:52-111
    import pymde_b200 as pm
    pm.seed(0)
    Y = torch.randn(10, 5)
    anchors = torch.tensor([0, 1, 3], device="cuda")
    values = torch.tensor([2.0, 1.0, 3.0], device="cuda").reshape(3, 1)
    c = pm.Anchored(anchors, values)
    for init in ("random", "quadratic"):
        mde = pm.preserve_neighbors(Y, embedding_dim=1, constraint=c, init=init)
        np.testing.assert_allclose(mde._X_init[anchors].cpu().numpy(), values.cpu().numpy())
    pm.seed(0)
    Y = torch.randn(3, 2)
    c = pm.Anchored(torch.tensor([0, 1], device="cuda"), torch.tensor([2.0, 3.0], device="cuda").reshape(2, 1))
    mde = pm.preserve_distances(Y, embedding_dim=1, constraint=c)
    assert mde.edges.tolist() == [[0, 2], [1, 2]]
    mde = pm.preserve_neighbors(Y, embedding_dim=1, constraint=c)
    assert mde.edges.tolist() == [[0, 2], [1, 2]]
    E = mde.embed(max_iter=20)
    np.testing.assert_allclose(E[:2].cpu().numpy(), [[2.0], [3.0]])


@pytest.mark.parametrize("n_items", [36, 1001])
def test_distances_reproducibility(n_items):
    # pymde/test_recipes.py:137-159: same seed => bit-identical edges and deviations
    import pymde_b200 as pm
    torch.manual_seed(0)
    Y = torch.rand((n_items, 128))
    prev = None
    for _ in range(3):
        pm.seed(0)
        mde = pm.preserve_distances(Y, max_distances=1e5)
        cur = (mde.edges.clone(), mde.distortion_function.deviations.clone())
        if prev is not None:
            assert torch.equal(cur[0], prev[0]) and torch.equal(cur[1], prev[1])
        prev = cur
