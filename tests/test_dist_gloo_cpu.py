"""CPU, world_size 2 over gloo: the edge-sharded evaluation scheme of pymde_b200/dist.py.

Each rank evaluates ITS edge range with the GLOBAL edge count as divisor (here with the numpy oracle
standing in for the CUDA kernel -- this test checks the host-side protocol, not the kernel), packs
[gradient | loss_hi | loss_lo] exactly like mde_solver.cu::pack_loss_kernel, sums the buffer with one
all-reduce, and must recover the unsharded value and gradient on every rank, bit-identically."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, REPO)
    from oracle import mde_oracle as O
    from pymde_b200.dist import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)  # same problem on every rank
    n, m, p = 300, 2, 4001
    e = rng.integers(0, n, (p * 2, 2))
    e = e[e[:, 0] != e[:, 1]][:p]
    w = rng.choice([1.0, 2.0, -1.0], p).astype(np.float32)
    X = rng.standard_normal((n, m)).astype(np.float32)
    lo, hi = shard_range(p, rank, world)
    spec = O.FnSpec(O.P_LOG1P, w[lo:hi], (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0))
    v, g = O.average_distortion(X.astype(np.float64), e[lo:hi], spec, True, np.float64, p_total=p)
    loss_sum = v * p  # what the kernel accumulates (not divided)
    buf = np.zeros(n * m + 2, dtype=np.float32)
    buf[: n * m] = g.astype(np.float32).ravel()
    buf[n * m] = np.float32(loss_sum)
    buf[n * m + 1] = np.float32(loss_sum - np.float64(np.float32(loss_sum)))
    t = torch.from_numpy(buf)
    dist.all_reduce(t)
    total = (float(t[n * m]) + float(t[n * m + 1])) / p
    full = O.FnSpec(O.P_LOG1P, w, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0))
    v_ref, g_ref = O.average_distortion(X.astype(np.float64), e, full, True, np.float64)
    ok = abs(total - v_ref) <= 1e-6 * abs(v_ref) and np.allclose(t[: n * m].numpy().reshape(n, m), g_ref, atol=1e-6)
    # every rank must hold the same bits after the all-reduce (replicated L-BFGS stays in lock-step)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    same = all(torch.equal(gathered[0], x) for x in gathered)
    ret[rank] = bool(ok and same)
    dist.destroy_process_group()


def test_sharded_evaluation_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r, False) for r in range(world)), dict(ret)


def _exchange_worker(rank, world, port, ret):
    sys.path.insert(0, REPO)
    from pymde_b200 import dist as pdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([(rank * 37 + k) % 256 for k in range(64)])
    got = pdist.make_exchange()(mine)  # what DeviceSolver hands to mde_solver_comm_connect
    want = b"".join(bytes([(r * 37 + k) % 256 for k in range(64)]) for r in range(world))
    ret[rank] = (got == want)
    dist.destroy_process_group()


def test_ipc_handle_exchange_world2():
    """The 64-byte cudaIpc handles reach every rank in rank order (host side of the peer-memory all-reduce)."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_exchange_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r, False) for r in range(world)), dict(ret)


def test_shard_ranges_cover_and_balance():
    sys.path.insert(0, REPO)
    from pymde_b200.dist import shard_range, pack_handles
    for p in (1, 7, 1000, 1554550):
        for world in (1, 2, 3, 8):
            spans = [shard_range(p, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == p
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pack_handles([b"short"])
