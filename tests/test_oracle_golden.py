"""CPU: the numpy oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py) and against the reference's own known-answer tests."""
import numpy as np
import pytest

from oracle import mde_oracle as O
from tests.golden_cases import CASES, spec_for


@pytest.mark.parametrize("name", sorted(CASES))
def test_function_values_and_derivatives(golden, name):
    g = golden["functions"]
    for tag, dtype, rtol in (("f64", np.float64, 1e-11), ("f32", np.float32, 2e-5)):
        spec = spec_for(name, g, tag)
        d = g["%s/%s/d" % (name, tag)]
        f, fp = O.eval_function(spec, d, dtype)
        np.testing.assert_allclose(f, g["%s/%s/f" % (name, tag)], rtol=rtol, atol=rtol)
        np.testing.assert_allclose(fp, g["%s/%s/fp" % (name, tag)], rtol=rtol, atol=rtol)


@pytest.mark.parametrize("key", ["m1", "m2", "m3", "m4", "m7", "m16", "m2_zero", "m3_zero"])
def test_average_distortion_value_and_grad(golden, key):
    g, fg = golden["evals"], golden["functions"]
    edges, X = g[key + "/edges"], g[key + "/X"]
    for name in sorted(CASES):
        for tag, dtype, rtol in (("f64", np.float64, 1e-10), ("f32", np.float32, 3e-5)):
            spec = spec_for(name, fg, tag)
            v, grad = O.average_distortion(X.astype(dtype), edges, spec, True, dtype)
            rv, rg = g["%s/%s/%s/value" % (key, name, tag)], g["%s/%s/%s/grad" % (key, name, tag)]
            np.testing.assert_allclose(v, rv, rtol=rtol, atol=rtol, equal_nan=True, err_msg=name)
            scale = max(1.0, float(np.abs(rg[np.isfinite(rg)]).max()) if np.isfinite(rg).any() else 1.0)
            np.testing.assert_allclose(grad, rg, rtol=rtol, atol=rtol * scale, equal_nan=True,
                                       err_msg=name)


def test_distances_golden(golden):
    g = golden["evals"]
    d, _ = O.edge_distances(g["m2/X"], g["m2/edges"])
    np.testing.assert_allclose(d, g["m2/pen_quadratic/distances"], rtol=1e-6)


def test_known_answer_62_over_3():
    # reference: pymde/test_optim.py:75-93
    edges = np.array([(0, 1), (0, 2), (1, 2)])
    X = np.array([[0.0, 0.0], [1.0, 1.0], [3.0, 3.0]])
    spec = O.FnSpec(O.P_QUADRATIC, np.array([1.0, 2.0, 3.0]))
    v, _ = O.average_distortion(X, edges, spec, False)
    np.testing.assert_allclose(v, 62.0 / 3, rtol=1e-12)


def test_grad_vs_dense_incidence():
    # reference: pymde/test_optim.py:97-118 (oracle: util.py:425-451, A diag(g) A^T X)
    rng = np.random.default_rng(0)
    edges = np.array([(0, 1), (0, 2), (1, 2)])
    X = rng.standard_normal((3, 2))
    w = np.array([1.0, 2.0, 3.0])
    _, grad = O.average_distortion(X, edges, O.FnSpec(O.P_QUADRATIC, w), True)
    A = np.array([[1, 1, 0], [-1, 0, 1], [0, -1, -1]], dtype=float)
    d = np.linalg.norm(A.T @ X, axis=1)
    gk = (2 * w * d / 3) / d
    np.testing.assert_allclose(grad, A @ (np.diag(gk) @ (A.T @ X)), rtol=1e-12)


def test_zero_distance_gives_zero_gradient():
    # reference: pymde/test_optim.py:57-71
    X = np.ones((3, 3))
    for fn in (O.P_QUADRATIC, O.P_LINEAR, O.P_LOG1P):
        _, grad = O.average_distortion(X, np.array([(0, 1)]), O.FnSpec(fn, np.ones(1), (1.5, 0, 0)))
        assert np.all(grad == 0.0)


@pytest.mark.parametrize("key", ["n2_m2", "n10_m3", "n100_m3", "n1000_m2", "n257_m5", "n300_m40"])
def test_projections(golden, key):
    g = golden["projections"]
    Z, G = g[key + "/Z"], g[key + "/G"]
    np.testing.assert_allclose(O.Centered().project(Z), g[key + "/centered"], atol=2e-6)
    Xs = O.Standardized().project(Z)
    np.testing.assert_allclose(Xs, g[key + "/standardized"], atol=2e-4, rtol=1e-4)
    n, m = Z.shape
    # invariants of pymde/test_util.py:20-71
    np.testing.assert_allclose(Xs.T @ Xs / n, np.eye(m), atol=1e-4)
    if n > 2:
        np.testing.assert_allclose(Xs.mean(0), 0, atol=1e-5)
    T = O.Standardized().tangent(g[key + "/standardized"], G)
    np.testing.assert_allclose(T, g[key + "/tangent"], atol=5e-5, rtol=1e-4)


def _spec_for_traj(key, par0):
    return {
        "quad_std": (O.FnSpec(O.P_QUADRATIC, par0), O.Standardized()),
        "pp_cen": (O.FnSpec(O.P_LOG1P, par0, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0)), O.Centered()),
        "pp_std": (O.FnSpec(O.P_LOG1P, par0, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0)), O.Standardized()),
        "cycle_abs": (O.FnSpec(O.L_ABSOLUTE, par0), O.Centered()),
        "huber_std": (O.FnSpec(O.L_HUBER, par0, (0.5, 0, 0)), O.Standardized()),
        "docs5": (O.FnSpec(O.P_QUADRATIC, par0), O.Standardized()),
    }[key]


TRAJ = ["quad_std", "pp_cen", "pp_std", "cycle_abs", "huber_std", "docs5"]


@pytest.mark.parametrize("key", TRAJ)
def test_trajectory_f64_matches_reference_f64(golden, key):
    """Pins the ALGORITHM (two-loop L-BFGS, strong-Wolfe search, projections, stale-gradient
    quirk, cached loss): the float64 restatement follows the reference's own float64 run of
    embed() iteration for iteration."""
    g = golden["trajectories"]
    spec, cons = _spec_for_traj(key, g[key + "/par0"].astype(np.float64))
    X, st = O.embed(g[key + "/X0"].astype(np.float64), g[key + "/edges"], spec, cons,
                    eps=float(g[key + "/eps"]), max_iter=int(g[key + "/max_iter"]), dtype=np.float64)
    ref = g[key + "/f64/average_distortions"]
    assert abs(st.iterations - len(ref)) <= 1  # convergence test can sit on a knife edge
    k = min(st.iterations, len(ref))
    np.testing.assert_allclose(st.average_distortions[:k], ref[:k], rtol=1e-6)
    np.testing.assert_allclose(st.residual_norms[:k], g[key + "/f64/residual_norms"][:k],
                               rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(st.step_size_percents[:k], g[key + "/f64/step_size_percents"][:k],
                               rtol=1e-4, atol=1e-9)
    if st.iterations == len(ref):
        np.testing.assert_allclose(X, g[key + "/f64/X"], atol=1e-4)


@pytest.mark.parametrize("key", TRAJ)
def test_trajectory_f32_close_to_reference_f32(golden, key):
    """fp32 runs differ from the reference only by summation-order noise: the first iterations
    agree closely, the final average distortion within the reference's own run-to-run spread
    (2e-4 .. 4e-3 relative after 300 iterations)."""
    g = golden["trajectories"]
    spec, cons = _spec_for_traj(key, g[key + "/par0"])
    X, st = O.embed(g[key + "/X0"], g[key + "/edges"], spec, cons, eps=float(g[key + "/eps"]),
                    max_iter=int(g[key + "/max_iter"]), dtype=np.float32)
    ref = g[key + "/average_distortions"]
    k = min(5, len(ref), st.iterations)
    np.testing.assert_allclose(st.average_distortions[:k], ref[:k], rtol=1e-3)
    np.testing.assert_allclose(st.average_distortions[0], ref[0], rtol=2e-6)
    final, _ = O.average_distortion(X, g[key + "/edges"], spec, False, np.float32)
    np.testing.assert_allclose(final, g[key + "/final_value"], rtol=1e-2)


@pytest.mark.parametrize("name", sorted(__import__("tests.golden_cases", fromlist=["CASES"]).CASES))
def test_closed_form_derivative_matches_central_differences(name):
    """Independent of the reference fixtures: the oracle's closed-form f' is the derivative of its f
    (float64 central differences, away from the kinks of the piecewise functions)."""
    from tests.golden_cases import CASES
    fa, sa, fr, sr = CASES[name]
    rng = np.random.default_rng(abs(hash_name(name)))
    p = 400
    d = rng.uniform(0.15, 3.0, p)
    if fr is not None:
        par0 = rng.choice([-1.0, 1.0, 2.0], p)
    elif fa >= O.L_ABSOLUTE:
        par0 = rng.uniform(0.3, 2.5, p)  # deviations
    else:
        par0 = rng.uniform(0.5, 2.0, p)  # weights
    par1 = rng.uniform(0.5, 2.0, p) if fa == O.L_WEIGHTED_QUADRATIC else None
    spec = O.FnSpec(fa, par0, sa, fn_rep=fr, rep=sr, par1=par1)
    h = 1e-6
    f, fp = O.eval_function(spec, d)
    fplus, _ = O.eval_function(spec, d + h)
    fminus, _ = O.eval_function(spec, d - h)
    num = (fplus - fminus) / (2 * h)
    # skip points within 1e-3 of a kink (|d - delta|, Huber thresholds, fractional d = delta)
    smooth = np.ones(p, dtype=bool)
    if fa >= O.L_ABSOLUTE:
        smooth &= np.abs(d - par0) > 1e-3
        if fa == O.L_HUBER:
            smooth &= np.abs(np.abs(d - par0) - sa[0]) > 1e-3
    if fa == O.P_HUBER:
        smooth &= np.abs(d - sa[0]) > 1e-3
    assert smooth.sum() > p // 2
    np.testing.assert_allclose(fp[smooth], num[smooth], rtol=2e-5, atol=1e-6)


def hash_name(name):
    import zlib
    return zlib.crc32(name.encode())


@pytest.mark.parametrize("m", [1, 2, 3, 5])
@pytest.mark.parametrize("name", ["pen_pushpull_log1p_log", "loss_huber_0.7", "pen_quadratic", "loss_soft_fractional_10"])
def test_gradient_of_average_distortion_matches_central_differences(name, m):
    """dE/dX of the oracle (closed form through f') against float64 central differences of E."""
    from tests.golden_cases import CASES
    fa, sa, fr, sr = CASES[name]
    rng = np.random.default_rng(hash_name(name) % 9973 + m)
    n, p = 12, 40
    e = rng.integers(0, n, (p, 2))
    e = e[e[:, 0] != e[:, 1]]
    p = len(e)
    par0 = rng.choice([-1.0, 1.0, 2.0], p) if fr is not None else rng.uniform(0.5, 2.0, p)
    spec = O.FnSpec(fa, par0, sa, fn_rep=fr, rep=sr)
    X = rng.standard_normal((n, m)) * 1.5
    v, g = O.average_distortion(X, e, spec, True)
    num = np.zeros_like(X)
    h = 1e-6
    for i in range(n):
        for c in range(m):
            Xp, Xm = X.copy(), X.copy()
            Xp[i, c] += h
            Xm[i, c] -= h
            num[i, c] = (O.average_distortion(Xp, e, spec, False)[0] - O.average_distortion(Xm, e, spec, False)[0]) / (2 * h)
    np.testing.assert_allclose(g, num, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(g.sum(0), 0.0, atol=1e-12)  # translation invariance


@pytest.mark.parametrize("key", ["quad", "pp"])
def test_anchored_trajectory_f64_matches_reference_f64(golden, key):
    """Anchored constraint (pymde/constraints.py:114-164): tangent projection zeroes the anchor rows, the retraction
    re-writes them; the float64 restatement follows the reference's float64 embed() iteration for iteration."""
    g = golden["anchored"]
    par0 = g[key + "/par0"].astype(np.float64)
    spec = (O.FnSpec(O.P_QUADRATIC, par0) if key == "quad"
            else O.FnSpec(O.P_LOG1P, par0, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0)))
    cons = O.Anchored(g["anchors"], g["values"].astype(np.float64))
    X, st = O.embed(g[key + "/X0"].astype(np.float64), g[key + "/edges"], spec, cons, eps=1e-6,
                    max_iter=int(g[key + "/max_iter"]), dtype=np.float64)
    ref = g[key + "/f64/average_distortions"]
    assert abs(st.iterations - len(ref)) <= 1
    k = min(st.iterations, len(ref))
    np.testing.assert_allclose(st.average_distortions[:k], ref[:k], rtol=1e-6)
    np.testing.assert_allclose(st.residual_norms[:k], g[key + "/f64/residual_norms"][:k], rtol=1e-4, atol=1e-9)
    np.testing.assert_array_equal(X[g["anchors"]], g["values"].astype(np.float64))
    if st.iterations == len(ref):
        np.testing.assert_allclose(X, g[key + "/f64/X"], atol=1e-4)
