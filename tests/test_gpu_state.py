"""Device copies must follow the Python objects they were taken from: the edge layout snapshots the
distortion function's parameters, the device solver the constraint's anchors; mutating or replacing either between
calls must change the result exactly as it does in the reference, which re-reads them at every evaluation
(pymde/average_distortion.py:47, pymde/constraints.py:150-164).  Also: the per-edge coefficient path (external
callables) for edge counts that are not a multiple of 4, in both layouts."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _problem(pm, n=400, p=3001, seed=0, m=2):
    rng = np.random.default_rng(seed)
    e = rng.integers(0, n, (2 * p, 2))
    e = e[e[:, 0] != e[:, 1]]
    e = np.unique(np.sort(e, axis=1), axis=0)[:p]
    X = rng.standard_normal((n, m)).astype(np.float32)
    return e, X


def test_reweighting_between_calls_is_seen():
    import pymde_b200 as pm
    from oracle import mde_oracle as O
    dev = torch.device("cuda", 0)
    e, X = _problem(pm)
    w = torch.ones(len(e), device=dev)
    f = pm.penalties.Quadratic(w)
    mde = pm.MDE(400, 2, torch.tensor(e, device=dev), f, pm.Centered(), device=dev)
    Xd = torch.tensor(X, device=dev)
    v1 = mde.average_distortion(Xd).item()
    f.weights *= 3.0  # in place
    v2 = mde.average_distortion(Xd).item()
    np.testing.assert_allclose(v2, 3.0 * v1, rtol=1e-6)
    f.weights = torch.full((len(e),), 0.5, device=dev)  # replaced
    v3 = mde.average_distortion(Xd).item()
    np.testing.assert_allclose(v3, 0.5 * v1, rtol=1e-6)
    mde.distortion_function = pm.losses.Absolute(torch.full((len(e),), 1.0, device=dev))  # another function object
    v4 = mde.average_distortion(Xd).item()
    ref, _ = O.average_distortion(X.astype(np.float64), e, O.FnSpec(O.L_ABSOLUTE, np.ones(len(e), np.float32)), False)
    np.testing.assert_allclose(v4, ref, rtol=1e-5)
    # the solver follows too: embed after the swap optimises the new objective
    mde.embed(X=Xd, max_iter=5)
    np.testing.assert_allclose(mde.solve_stats.average_distortions[0], ref, rtol=1e-5)


def test_anchor_values_mutated_between_embeds():
    import pymde_b200 as pm
    dev = torch.device("cuda", 0)
    e, X = _problem(pm)
    anchors = torch.arange(10, device=dev)
    values = torch.zeros(10, 2, device=dev)
    cons = pm.Anchored(anchors, values)
    mde = pm.MDE(400, 2, torch.tensor(e, device=dev), pm.penalties.Quadratic(torch.ones(len(e), device=dev)), cons,
                 device=dev)
    X1 = mde.embed(max_iter=10)
    assert float(X1[:10].abs().max()) == 0.0
    cons.values += 2.0  # in place: the cached solver must not keep the old anchor rows
    X2 = mde.embed(max_iter=10)
    assert torch.equal(X2[:10], torch.full((10, 2), 2.0, device=dev))
    cons2 = pm.Anchored(anchors, torch.full((10, 2), -1.0, device=dev))
    mde.constraint = cons2
    X3 = mde.embed(max_iter=10)
    assert torch.equal(X3[:10], torch.full((10, 2), -1.0, device=dev))


def test_push_and_pull_with_non_table_penalty():
    """PushAndPull accepts any callable penalty class (reference penalties.py:384-400): Hinge / Sigmoid have no
    kernel id, the object must still evaluate (masked torch path) and embed (external-callable path)."""
    import pymde_b200 as pm
    dev = torch.device("cuda", 0)
    e, X = _problem(pm)
    w = torch.tensor(np.where(np.arange(len(e)) % 2 == 0, 1.0, -1.0).astype(np.float32), device=dev)
    import functools
    hinge = functools.partial(pm.penalties.Hinge, threshold=1.0)  # the recipes pass penalty classes the same way
    f = pm.penalties.PushAndPull(w, pm.penalties.Log1p, hinge)
    d = torch.rand(len(e), device=dev) + 0.1
    out = f(d)
    pos = w >= 0
    np.testing.assert_allclose(out[pos].cpu().numpy(), np.log1p(d[pos].cpu().numpy() ** 1.5), rtol=1e-5)
    assert torch.equal(out[~pos], hinge(w[~pos])(d[~pos]))
    mde = pm.MDE(400, 2, torch.tensor(e, device=dev), f, pm.Centered(), device=dev)
    mde.embed(X=torch.tensor(X, device=dev), max_iter=8)
    st = mde.solve_stats
    assert st.iterations == 8 and st.average_distortions[-1] <= st.average_distortions[0]


@pytest.mark.parametrize("layout", ["soa", "tiles", "pull"])
@pytest.mark.parametrize("extra", [1, 2, 3])
def test_external_coefficients_edge_count_not_multiple_of_4(layout, extra, monkeypatch):
    """mde_scatter_external reads the permutation four entries at a time (MODE 2 of the quad / tile kernels)."""
    import pymde_b200 as pm
    dev = torch.device("cuda", 0)
    monkeypatch.setenv("MDE_B200_LAYOUT", layout)
    e, X = _problem(pm, p=2000 + extra, seed=extra)
    assert len(e) % 4 == extra % 4
    wts = torch.rand(len(e), device=dev) + 0.5
    mde = pm.MDE(400, 2, torch.tensor(e, device=dev), lambda d: wts * d ** 2, pm.Centered(), device=dev)
    Xd = torch.tensor(X, device=dev, requires_grad=True)
    v = mde.average_distortion(Xd)
    v.backward()
    Xr = torch.tensor(X, device=dev, requires_grad=True)
    et = torch.tensor(e, device=dev)
    ref = (wts * (Xr[et[:, 0]] - Xr[et[:, 1]]).pow(2).sum(1)).mean()
    ref.backward()
    np.testing.assert_allclose(v.item(), ref.item(), rtol=1e-5)
    np.testing.assert_allclose(Xd.grad.cpu().numpy(), Xr.grad.cpu().numpy(), atol=2e-5 * float(Xr.grad.abs().max()))


@pytest.mark.parametrize("m", [1, 2, 3, 4])
def test_deterministic_mode_is_bit_reproducible(m, monkeypatch):
    """MDE_B200_DETERMINISTIC=1: 64-bit fixed-point accumulation of the gradient contributions (integer adds are
    associative), fixed-order loss sums => value, gradient and whole embed() trajectories are bit-identical run to run.
    The reference's scatter_add_ on CUDA is not (pymde/average_distortion.py:75-76)."""
    import pymde_b200 as pm
    from pymde_b200 import _lib
    from oracle import mde_oracle as O
    dev = torch.device("cuda", 0)
    monkeypatch.setenv("MDE_B200_DETERMINISTIC", "1")
    rng = np.random.default_rng(10 + m)
    n, p = 3000, 60000
    e = rng.integers(0, n, (2 * p, 2))
    e = np.unique(np.sort(e[e[:, 0] != e[:, 1]], axis=1), axis=0)[:p]
    w = rng.choice([1.0, 2.0, -1.0], len(e)).astype(np.float32)
    X0 = rng.standard_normal((n, m)).astype(np.float32)
    X0 -= X0.mean(0)

    def problem():
        f = pm.penalties.PushAndPull(torch.tensor(w, device=dev), pm.penalties.Log1p, pm.penalties.Log)
        return pm.MDE(n, m, torch.tensor(e, device=dev), f, pm.Centered(), device=dev)

    grads, values = [], []
    for _ in range(3):
        mde = problem()
        assert _lib.load().mde_edges_deterministic(mde._layout().handle) == 1
        X = torch.tensor(X0, device=dev, requires_grad=True)
        v = mde.average_distortion(X)
        v.backward()
        grads.append(X.grad.clone())
        values.append(v.item())
    assert values[0] == values[1] == values[2]
    assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
    spec = O.FnSpec(O.P_LOG1P, w, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0))
    v_ref, g_ref = O.average_distortion(X0.astype(np.float64), e, spec, True)
    np.testing.assert_allclose(values[0], v_ref, rtol=1e-5)
    np.testing.assert_allclose(grads[0].cpu().numpy(), g_ref, atol=3e-5 * np.abs(g_ref).max())
    runs = []
    for _ in range(2):
        mde = problem()
        runs.append((mde.embed(X=torch.tensor(X0, device=dev), max_iter=40, eps=0.0).clone(),
                     list(mde.solve_stats.average_distortions)))
    assert torch.equal(runs[0][0], runs[1][0])
    assert runs[0][1] == runs[1][1]
