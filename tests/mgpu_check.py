#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun): the edge-sharded solve must reproduce the single-GPU solve.

Every rank builds its shard of the SAME problem (pymde_b200.dist.shard_mde), embeds K iterations from the same
X0; rank 0 also solves the unsharded problem.  Checks: (i) per-iteration losses agree (first iterations to 1e-5:
the sums only differ in association order), (ii) all ranks end with bit-identical X (replicated L-BFGS stays in
lock-step), (iii) the sharded value of a fixed X equals the oracle's."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
import bench
import pymde_b200 as pm
from pymde_b200 import dist as pdist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
td.init_process_group("nccl", device_id=dev)
n, m = 20000, 2
edges, w = bench.c2_edges(0, n=n, k=10)
X0 = bench.initial_iterate(0, n=n, m=m)
et, wt = torch.tensor(edges), torch.tensor(w)
mk = lambda lo, hi: pm.penalties.PushAndPull(wt[lo:hi].to(dev), pm.penalties.Log1p, pm.penalties.Log)
mde = pdist.shard_mde(pm.MDE, n, m, et, mk, pm.Centered(), dev)
K = 25
X = mde.embed(X=torch.tensor(X0, device=dev), max_iter=K, eps=0.0)
st = mde.solve_stats
# (ii) identical X on every rank
chk = X.double().sum().reshape(1); allc = [torch.zeros_like(chk) for _ in range(world)]
td.all_gather(allc, chk)
same = all(float(c) == float(allc[0]) for c in allc)
ok = True
if rank == 0:
    full = pm.MDE(n, m, et.to(dev), mk(0, len(edges)), pm.Centered(), device=dev)
    Xf = full.embed(X=torch.tensor(X0, device=dev), max_iter=K, eps=0.0)
    a, b = np.array(st.average_distortions), np.array(full.solve_stats.average_distortions)
    ra, rb = st.residual_norms[0], full.solve_stats.residual_norms[0]
    from oracle import mde_oracle as O
    spec = O.FnSpec(O.P_LOG1P, w, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0))
    v_ref, g_ref = O.average_distortion(X0.astype(np.float64), edges, spec, True)
    r_ref = float(np.sqrt((g_ref ** 2).sum()))
    print("sharded x%d: iters %d/%d | loss[0] sharded %.8f full %.8f oracle %.8f | ||grad||[0] sharded %.8g full %.8g "
          "oracle %.8g | final %.6f vs %.6f (non-converged runs are chaotic: +-15%% run to run on ONE gpu) | "
          "X identical across ranks: %s" % (world, st.iterations, full.solve_stats.iterations, a[0], b[0], v_ref,
                                            ra, rb, r_ref, a[-1], b[-1], same))
    ok = (same and abs(a[0] - v_ref) < 1e-5 * abs(v_ref) and abs(ra - r_ref) < 1e-4 * r_ref
          and a[-1] < 0.5 * a[0] and abs(a[-1] - b[-1]) < 0.3 * abs(b[-1]))
    print("MGPU_CHECK", "OK" if ok else "FAILED")
td.barrier()
td.destroy_process_group()
sys.exit(0 if ok else 1)
