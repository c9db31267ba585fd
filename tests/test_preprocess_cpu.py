"""CPU: problem-construction helpers (host logic around the hot path): Graph semantics
(pymde/preprocess/test_graph.py), edge sampling, shortest paths, scaling, edge sharding."""
import numpy as np
import pytest
import torch

from pymde_b200.preprocess import preprocess as P
from pymde_b200.preprocess.graph import Graph, k_nearest_neighbors, shortest_paths
from pymde_b200.dist import shard_range


def test_graph_from_edges_sums_duplicates_and_sorts():
    # reference: pymde/preprocess/test_graph.py (duplicate / reciprocal edges are summed; i < j sorted)
    e = np.array([[0, 1], [1, 0], [1, 2], [2, 3], [3, 0]])
    g = Graph.from_edges(e)
    assert g.edges.tolist() == [[0, 1], [0, 3], [1, 2], [2, 3]]
    assert g.weights.tolist() == [2.0, 1.0, 1.0, 1.0]
    assert g.n_items == 4 and g.n_edges == 4 and g.n_all_edges == 6
    assert g.edges.dtype == torch.int64 and g.weights.dtype == torch.float32
    assert sorted(g.neighbors(0).tolist()) == [1, 3]
    with pytest.raises(ValueError):
        Graph(np.eye(3))


def test_shortest_paths_cycle_1000():
    # reference: pymde/preprocess/test_graph.py:125-162 (1000-node cycle)
    n = 1000
    g = Graph.from_edges(np.array([(i, (i + 1) % n) for i in range(n)]))
    d = shortest_paths(g)
    assert d.n_edges == n * (n - 1) // 2
    e, v = d.edges.numpy(), d.distances.numpy()
    hop = np.minimum(e[:, 1] - e[:, 0], n - (e[:, 1] - e[:, 0]))
    np.testing.assert_array_equal(v, hop.astype(np.float32))
    lim = shortest_paths(g, max_length=3)
    assert float(lim.distances.max()) == 3.0 and lim.n_edges == 3 * n


def test_graph_knn_on_cycle():
    n = 200
    g = Graph.from_edges(np.array([(i, (i + 1) % n) for i in range(n)]))
    kn = k_nearest_neighbors(g, 4)
    assert kn.n_edges == 2 * n  # neighbours at hop 1 and 2, each pair reciprocal => weight 2
    assert kn.weights.unique().tolist() == [2.0]


def test_sample_edges_excludes_and_dedups():
    n = 60
    excl = torch.tensor([[i, i + 1] for i in range(n - 1)])
    s = P.sample_edges(n, 500, exclude=excl, seed=0)
    assert s.shape[0] <= 500 and s.shape[0] > 450
    assert bool((s[:, 0] < s[:, 1]).all())
    keys = (s[:, 0] * n + s[:, 1]).numpy()
    assert len(np.unique(keys)) == len(keys)
    assert not np.isin(keys, (excl[:, 0] * n + excl[:, 1]).numpy()).any()
    s2 = P.sample_edges(n, 500, exclude=excl, seed=0)
    assert torch.equal(s, s2)  # same seed => identical edges (pymde/test_recipes.py:115-159)
    with pytest.raises(ValueError):
        P.sample_edges(5, 11)


def test_scale_and_dedup():
    d = torch.tensor([1.0, 2.0, 3.0])
    s = P.scale(d, 2.0)
    np.testing.assert_allclose(float(s.pow(2).mean().sqrt()), 2.0, rtol=1e-6)
    e = P.deduplicate_edges(torch.tensor([[1, 0], [0, 1], [2, 1]]))
    assert e.tolist() == [[0, 1], [1, 2]]


def test_shard_ranges_partition_the_edge_list():
    for p in (1, 7, 1000, 1554550):
        for w in (1, 2, 3, 8):
            r = [shard_range(p, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == p
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_spectral_initialisation_is_standardized_and_spectral():
    # pymde/quadratic.py:122-179: bottom non-trivial eigenvectors of the Laplacian, standardized
    from pymde_b200 import quadratic
    for n, m in ((3, 1), (10, 1), (5, 2), (40, 2), (300, 3)):
        e = torch.tensor([[i, (i + 1) % n] for i in range(n if n > 3 else n - 1)])
        X = quadratic.spectral(n, m, e, torch.ones(e.shape[0])).double().cpu()
        np.testing.assert_allclose((X.T @ X / n).numpy(), np.eye(m), atol=1e-5)
        assert float(X.mean(0).abs().max()) < 1e-5
    # a cycle embeds as a circle: sum of squared edge lengths equals n * 2 * (1 - cos(2 pi / n)) * ... (m = 2)
    n = 40
    e = torch.tensor([[i, (i + 1) % n] for i in range(n)])
    X = quadratic.spectral(n, 2, e, torch.ones(n)).double().cpu()
    d2 = (X[e[:, 0]] - X[e[:, 1]]).pow(2).sum(1)
    lam = 2.0 * (1.0 - np.cos(2.0 * np.pi / n))  # smallest non-zero Laplacian eigenvalue of the cycle (double)
    np.testing.assert_allclose(float(d2.sum()), 2.0 * n * lam, rtol=1e-4)


def test_shortest_paths_retain_fraction_and_graph_helpers():
    import pymde_b200 as pm
    from pymde_b200.preprocess import graph as G
    n = 300
    g = Graph.from_edges(np.array([(i, (i + 1) % n) for i in range(n)]))
    full = shortest_paths(g)
    pm.seed(0)
    a = shortest_paths(g, retain_fraction=0.25)
    pm.seed(0)
    b = shortest_paths(g, retain_fraction=0.25)
    assert torch.equal(a.edges, b.edges) and torch.equal(a.distances, b.distances)  # seeded => reproducible
    assert 0.2 * full.n_edges < a.n_edges < 0.3 * full.n_edges
    # every retained pair carries the same distance as in the full graph
    key = lambda e: (e[:, 0] * n + e[:, 1]).numpy()
    pos = np.searchsorted(key(full.edges), key(a.edges))
    np.testing.assert_array_equal(full.distances.numpy()[pos], a.distances.numpy())
    # scale(): RMS of the distances becomes the natural length, structure untouched
    sc = G.scale(full, 2.0)
    np.testing.assert_allclose(float(sc.distances.pow(2).mean().sqrt()), 2.0, rtol=1e-5)
    assert torch.equal(sc.edges, full.edges)
    # breadth_first_order(): hop counts with inf for unreachable nodes, -9999 for "no predecessor"
    h = Graph.from_edges(np.array([[0, 1], [1, 2], [3, 4]]), n_items=6)
    lengths, pred = G.breadth_first_order(h.A, 0)
    assert lengths.tolist() == [0.0, 1.0, 2.0, np.inf, np.inf, np.inf]
    assert pred[1] == 0 and pred[2] == 1 and pred[0] == -9999 and pred[5] == -9999
    assert pm.preprocess.generic.distances is pm.preprocess.distances
    with pytest.raises(ValueError):
        shortest_paths("not a graph")
