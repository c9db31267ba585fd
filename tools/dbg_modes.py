import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench, pymde_b200 as pm
from pymde_b200 import optim
n, m = 20000, 2
edges, w = bench.c2_edges(0, n=n, k=10); X0 = bench.initial_iterate(0, n=n, m=m)
dev = torch.device('cuda', 0)
for mode in (0, 1, 0, 1):
    optim.DEFAULT_MODE = mode
    f = pm.penalties.PushAndPull(torch.tensor(w, device=dev), pm.penalties.Log1p, pm.penalties.Log)
    mde = pm.MDE(n, m, torch.tensor(edges, device=dev), f, pm.Centered(), device=dev)
    mde.embed(X=torch.tensor(X0, device=dev), max_iter=25, eps=0.0)
    st = mde.solve_stats
    print('mode', mode, ['%.6f' % v for v in st.average_distortions[:6]], 'final %.6f' % st.average_distortions[-1], ['%.4g' % v for v in st.step_lengths[:6]], st.func_evals)
