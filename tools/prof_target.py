"""Profiling target (run under ncu): the C2-shaped bench problem, a few solver iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import pymde_b200 as pm
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cons = pm.Standardized() if (len(sys.argv) > 2 and sys.argv[2] == "std") else pm.Centered()
edges, w = bench.c2_edges(0)
X0 = bench.initial_iterate(0)
dev = torch.device("cuda", 0)
f = pm.penalties.PushAndPull(torch.tensor(w, device=dev), pm.penalties.Log1p, pm.penalties.Log)
mde = pm.MDE(bench.N_ITEMS, bench.EMBED_DIM, torch.tensor(edges, device=dev), f, cons, device=dev)
X = mde.embed(X=torch.tensor(X0, device=dev), max_iter=iters, eps=0.0)
torch.cuda.synchronize()
print("done", mde.solve_stats.iterations, mde.solve_stats.average_distortions[-1])
