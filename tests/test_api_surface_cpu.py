"""CPU: the package is a drop-in for the reference's public surface.  Signatures are compared with the
unmodified reference (loaded from baseline/_ref or /root/reference; skipped when neither is present), and the
host-side helpers that do not need the GPU are checked against it value for value."""
import inspect

import numpy as np
import pytest
import torch

import pymde_b200 as pm


@pytest.fixture(scope="module")
def ref():
    try:
        from oracle.ref_loader import load_reference
        return load_reference()
    except Exception as e:  # pragma: no cover
        pytest.skip("reference not importable here: %s" % e)


def _params(f):
    return [p for p in inspect.signature(f).parameters if p != "self"]


def test_public_signatures_match_the_reference(ref):
    pairs = {
        "MDE.__init__": (ref.MDE.__init__, pm.MDE.__init__),
        "MDE.embed": (ref.MDE.embed, pm.MDE.embed),
        "MDE.average_distortion": (ref.MDE.average_distortion, pm.MDE.average_distortion),
        "MDE.distortions": (ref.MDE.distortions, pm.MDE.distortions),
        "MDE.distances": (ref.MDE.distances, pm.MDE.distances),
        "MDE.differences": (ref.MDE.differences, pm.MDE.differences),
        "MDE.high_distortion_pairs": (ref.MDE.high_distortion_pairs, pm.MDE.high_distortion_pairs),
        "preserve_neighbors": (ref.preserve_neighbors, pm.preserve_neighbors),
        "preserve_distances": (ref.preserve_distances, pm.preserve_distances),
        "laplacian_embedding": (ref.laplacian_embedding, pm.laplacian_embedding),
        "Anchored": (ref.Anchored.__init__, pm.Anchored.__init__),
        "Graph.from_edges": (ref.Graph.from_edges, pm.Graph.from_edges),
        "quadratic.spectral": (ref.quadratic.spectral, pm.quadratic.spectral),
        "pca": (ref.pca, pm.pca),
        "align": (ref.align, pm.align),
        "rotate": (ref.rotate, pm.rotate),
        "util.proj_standardized": (ref.util.proj_standardized, pm.util.proj_standardized),
        "preprocess.dissimilar_edges": (ref.preprocess.dissimilar_edges, pm.preprocess.dissimilar_edges),
    }
    for cls in ("Linear", "Quadratic", "Cubic", "Power", "Huber", "Logistic", "Sigmoid", "Hinge", "Log1p", "Log",
                "InvPower", "LogRatio", "PushAndPull"):
        pairs["penalties." + cls] = (getattr(ref.penalties, cls).__init__, getattr(pm.penalties, cls).__init__)
    for cls in ("Absolute", "Quadratic", "WeightedQuadratic", "Huber", "Cubic", "Power", "Logistic", "Fractional",
                "SoftFractional"):
        pairs["losses." + cls] = (getattr(ref.losses, cls).__init__, getattr(pm.losses, cls).__init__)
    for name, (a, b) in pairs.items():
        assert _params(a) == _params(b), name
    # same leading parameters; ours may add a trailing `device=` (the data always ends up on a CUDA device)
    for name in ("sample_edges", "k_nearest_neighbors", "distances"):
        a, b = _params(getattr(ref.preprocess, name)), _params(getattr(pm.preprocess, name))
        assert b[: len(a)] == a and set(b[len(a):]) <= {"device"}, name


def test_public_names_exist(ref):
    out_of_scope = {"latexify", "plot", "datasets", "experiment_utils"}  # plotting / downloads
    for name in dir(ref):
        if name.startswith("_") or name in out_of_scope:
            continue
        obj = getattr(ref, name)
        if inspect.isfunction(obj) or inspect.isclass(obj):
            assert hasattr(pm, name), name
    assert hasattr(pm.Graph, "draw") and hasattr(pm.preprocess, "generic")
    for name in ("scale", "shortest_paths", "k_nearest_neighbors", "breadth_first_order", "Graph"):
        assert hasattr(pm.preprocess.graph, name), name
    assert _params(ref.preprocess.graph.shortest_paths) == _params(pm.preprocess.graph.shortest_paths)
    assert _params(ref.Graph.draw) == _params(pm.Graph.draw)
    for mod in ("penalties", "losses", "constraints", "preprocess", "recipes", "quadratic"):
        r, o = getattr(ref, mod), getattr(pm, mod)
        for name in dir(r):
            obj = getattr(r, name)
            if name.startswith("_") or not (inspect.isfunction(obj) or inspect.isclass(obj)):
                continue
            if getattr(obj, "__module__", "").startswith("pymde"):
                assert hasattr(o, name), "%s.%s" % (mod, name)


def test_host_helpers_agree_with_the_reference(ref):
    torch.manual_seed(0)
    w, d = torch.randn(64), torch.rand(64) * 3
    for name, args in (("Sigmoid", (1.0, 2.0)), ("Sigmoid", (0.5,)), ("Hinge", (1.0,)), ("Hinge", (1.5, 0.2))):
        a = getattr(ref.penalties, name)(w, *args)(d)
        b = getattr(pm.penalties, name)(w, *args)(d)
        assert torch.equal(a, b), name
    X2, X3 = torch.randn(9, 2), torch.randn(9, 3)
    np.testing.assert_allclose(pm.rotate(X2, 33.0).numpy(), ref.rotate(X2, torch.tensor(33.0)).numpy(), atol=1e-6)
    deg = torch.tensor([10.0, 20.0, 30.0])
    np.testing.assert_allclose(pm.rotate(X3, deg).numpy(), ref.rotate(X3, deg).numpy(), atol=1e-6)
    with pytest.raises(ValueError):
        pm.rotate(torch.randn(4, 5), 1.0)
    assert torch.equal(pm.util.random_edges(50, 100, seed=3), ref.util.random_edges(50, 100, seed=3))
    e, wt = torch.tensor([[0, 1], [1, 3]]), torch.tensor([1.0, 2.0])
    assert abs(pm.util.adjacency_matrix(5, 2, e, wt) - ref.util.adjacency_matrix(5, 2, e, wt)).max() == 0
    Xq = torch.tensor([[1.0, 1.0], [1.0, -1.0], [-1.0, 1.0], [-1.0, -1.0]])  # exactly standardized
    assert pm.util.in_stdemb(Xq) and bool(ref.util.in_stdemb(Xq)) and not pm.util.in_stdemb(1.1 * Xq)
    for seed in range(4):  # same (tolerance-sensitive) verdict as the reference on arbitrary inputs
        torch.manual_seed(seed)
        Xs = ref.util.proj_standardized(torch.randn(40, 2), demean=True)
        assert pm.util.in_stdemb(Xs) == bool(ref.util.in_stdemb(Xs))
    Y = torch.randn(30, 6)
    a, b = ref.pca(Y, 2), pm.pca(Y, 2)
    np.testing.assert_allclose((a.abs()).numpy(), (b.abs()).numpy(), rtol=1e-4, atol=1e-5)  # columns up to sign
    al = pm.align(a, b)
    np.testing.assert_allclose(al.numpy(), ref.align(a, b).numpy(), atol=1e-5)


def test_default_device_is_cuda_only():
    assert pm.util.get_default_device().startswith("cuda")
    with pytest.raises(ValueError):
        pm.util.set_default_device("cpu")
