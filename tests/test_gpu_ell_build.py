"""The ELL pull records as the library builds them on the DEVICE (mde_ell.cu::ell_build_device: radix sorts, scans,
one fill kernel) against the HOST builder that tests/test_ell_layout_cpu.py checks against the oracle: same bytes, same
tables.  Then both builders behind MDE(...) give the same evaluation."""
import ctypes as C

import numpy as np
import pytest
import torch

from pymde_b200 import _lib
from tests.test_ell_layout_cpu import build as host_build, random_problem

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
KEYS = ("rec", "rec_off", "bkt_tile", "bkt_wt0", "cta_wt0", "cta_bkt0")


def device_build(n, m, edges, w, push_pull, rb=0, max_cta=0):
    lib = _lib.load()
    src = torch.tensor(np.ascontiguousarray(edges[:, 0], dtype=np.int32), device=DEV)
    dst = torch.tensor(np.ascontiguousarray(edges[:, 1], dtype=np.int32), device=DEV)
    wd = torch.tensor(np.ascontiguousarray(w, dtype=np.float32), device=DEV)
    h = _lib.mde_ell_host_t()
    st = torch.cuda.current_stream(DEV).cuda_stream
    rc = lib.mde_ell_device_layout(n, src.numel(), m, src.data_ptr(), dst.data_ptr(), wd.data_ptr(), int(push_pull), rb,
                                   max_cta, C.byref(h), st)
    if rc != 0:
        return rc, None
    out = dict(
        rec=np.ctypeslib.as_array(h.rec, shape=(h.rec_bytes,)).copy(),
        rec_off=np.ctypeslib.as_array(h.rec_off, shape=(h.nrec + 1,)).copy(),
        bkt_tile=np.ctypeslib.as_array(h.bkt_tile, shape=(h.nbkt,)).copy(),
        bkt_wt0=np.ctypeslib.as_array(h.bkt_wt0, shape=(h.nbkt + 1,)).copy(),
        cta_wt0=np.ctypeslib.as_array(h.cta_wt0, shape=(h.ncta + 1,)).copy(),
        cta_bkt0=np.ctypeslib.as_array(h.cta_bkt0, shape=(h.ncta,)).copy(),
        nrec=h.nrec, nslots=h.nslots, nentries=h.nentries, npadded=h.npadded, rb=h.tile_rows_log2, ncta=h.ncta)
    lib.mde_ell_host_free(C.byref(h))
    return 0, out


def assert_same(a, b):
    for k in ("nrec", "nslots", "nentries", "npadded", "rb", "ncta"):
        assert a[k] == b[k], k
    for k in KEYS:
        assert a[k].shape == b[k].shape, k
        if not np.array_equal(a[k], b[k]):
            bad = np.flatnonzero(a[k] != b[k])
            raise AssertionError("%s differs at %d positions, first %s" % (k, len(bad), bad[:8]))


@pytest.mark.parametrize("m", [1, 2, 3, 4])
@pytest.mark.parametrize("push_pull", [False, True])
@pytest.mark.parametrize("n,p,rb,local", [(300, 2500, 8, False), (1000, 12000, 8, True), (64, 40, 0, False),
                                          (5000, 30000, 10, True), (40, 700, 8, False), (20000, 400000, 0, False)])
def test_device_builder_matches_host_builder_bit_for_bit(m, push_pull, n, p, rb, local):
    rng = np.random.default_rng(31 * m + n + push_pull)
    edges, w = random_problem(rng, n, p, push_pull, local)
    rc_h, host = host_build(n, m, edges, w, push_pull, rb)
    rc_d, dev = device_build(n, m, edges, w, push_pull, rb)
    assert rc_h == 0 and rc_d == 0
    assert_same(host, dev)


def test_device_builder_on_the_bench_workload():
    import bench
    edges, w = bench.c2_edges(0)
    e = np.sort(edges, axis=1)
    rc_h, host = host_build(bench.N_ITEMS, 2, e, w, True)
    rc_d, dev = device_build(bench.N_ITEMS, 2, e, w, True)
    assert rc_h == 0 and rc_d == 0
    assert_same(host, dev)


def test_unsupported_shapes_are_refused_on_the_device_too():
    rng = np.random.default_rng(6)
    edges, w = random_problem(rng, 9000, 2000, False, False)
    assert device_build(9000, 2, edges, w, False, rb=8)[0] == _lib.MDE_E_UNSUPPORTED


@pytest.mark.parametrize("builder", ["host", "device"])
def test_both_builders_behind_mde_give_the_same_evaluation(builder, monkeypatch):
    import pymde_b200 as pm
    monkeypatch.setenv("MDE_B200_LAYOUT", "ell")
    monkeypatch.setenv("MDE_B200_ELL_BUILD", builder)
    rng = np.random.default_rng(3)
    n, m = 6000, 2
    edges, w = random_problem(rng, n, 90000, True, True)
    e = torch.tensor(edges, device=DEV)
    f = pm.penalties.PushAndPull(torch.tensor(w, device=DEV), pm.penalties.Log1p, pm.penalties.Log)
    mde = pm.MDE(n, m, e, f, pm.Centered(), device=DEV)
    X = torch.tensor(rng.standard_normal((n, m)).astype(np.float32), device=DEV, requires_grad=True)
    v = mde.average_distortion(X)
    v.backward()
    assert _lib.load().mde_edges_kind(mde._layout().handle) == 3
    monkeypatch.setenv("MDE_B200_LAYOUT", "soa")
    ref = pm.MDE(n, m, e, f, pm.Centered(), device=DEV)
    X2 = X.detach().clone().requires_grad_(True)
    v2 = ref.average_distortion(X2)
    v2.backward()
    np.testing.assert_allclose(v.item(), v2.item(), rtol=1e-6)
    g, g2 = X.grad.cpu().numpy(), X2.grad.cpu().numpy()
    np.testing.assert_allclose(g, g2, atol=2e-6 * np.abs(g2).max())
