"""CPU: the C restatement (oracle/mde_oracle.c) against the pinned numpy oracle and the reference fixtures."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import mde_oracle as O
from tests.golden_cases import CASES, spec_for


@pytest.mark.parametrize("key", ["m1", "m2", "m3", "m7", "m16"])
def test_c_oracle_matches_reference_fixtures(golden, key):
    g, fg = golden["evals"], golden["functions"]
    edges, X = g[key + "/edges"], g[key + "/X"]
    for name in sorted(CASES):
        spec = spec_for(name, fg, "f32")
        v, grad = c_oracle.average_distortion(X, edges, spec)
        rv, rg = g["%s/%s/f64/value" % (key, name)], g["%s/%s/f64/grad" % (key, name)]
        np.testing.assert_allclose(v, rv, rtol=2e-6, err_msg=name)
        np.testing.assert_allclose(grad, rg, rtol=1e-5, atol=2e-6 * max(1.0, np.abs(rg).max()), err_msg=name)


def test_c_oracle_matches_numpy_oracle_large():
    rng = np.random.default_rng(1)
    n, m, p = 5000, 3, 200000
    e = rng.integers(0, n, (p, 2))
    e = e[e[:, 0] != e[:, 1]]
    w = rng.choice([1.0, 2.0, -1.0], len(e)).astype(np.float32)
    X = rng.standard_normal((n, m)).astype(np.float32)
    spec = O.FnSpec(O.P_LOG1P, w, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0))
    v1, g1 = O.average_distortion(X.astype(np.float64), e, spec, True, np.float64)
    v2, g2 = c_oracle.average_distortion(X, e, spec)
    np.testing.assert_allclose(v2, v1, rtol=1e-12)
    np.testing.assert_allclose(g2, g1, rtol=1e-9, atol=1e-14)
