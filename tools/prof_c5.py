"""Profiling target (run under ncu): the fused distortion kernel on one C5 shard (n = 1e7, 2.5e7 SBM edges, m = 2),
layout from MDE_B200_LAYOUT (default: the library's choice).  A few launches, inputs larger than L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pymde_b200 as pm
from pymde_b200 import _lib
dev = torch.device("cuda", 0)
p = int(sys.argv[1]) if len(sys.argv) > 1 else bench.C5_SHARD_EDGES
edges, w = bench.c5_shard(0, p=p)
wt = torch.tensor(w, device=dev)
mde = pm.MDE(bench.C5_N, 2, torch.tensor(edges, device=dev), pm.penalties.PushAndPull(wt, pm.penalties.Log1p, pm.penalties.Log),
             pm.Centered(), device=dev)
X = torch.tensor(bench.initial_iterate(2, bench.C5_N, 2), device=dev)
lay = mde._layout()
lib = _lib.load()
g = torch.zeros_like(X)
st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(4):
    g.zero_()
    _lib.check(lib.mde_distortion(lay.handle, X.data_ptr(), 2, g.data_ptr(), None, st))
torch.cuda.synchronize()
print("kind", lib.mde_edges_kind(lay.handle), "edges", len(edges))
