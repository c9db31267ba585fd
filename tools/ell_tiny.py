"""Small ELL-layout evaluation against the sorted-SoA kernels (run under compute-sanitizer when a parity test fails)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pymde_b200 as pm
from pymde_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
for n, m, rb in ((3000, 2, "9"), (700, 3, "8"), (5000, 2, ""), (900, 1, "8"), (1200, 4, "8")):
    i = np.repeat(np.arange(n), 9)
    j = (i + rng.integers(1, 60, len(i))) % n
    att = np.unique(np.sort(np.stack([i, j], 1), axis=1), axis=0)
    rep = rng.integers(0, n, (len(att), 2))
    rep = np.unique(np.sort(rep[rep[:, 0] != rep[:, 1]], axis=1), axis=0)
    edges = torch.tensor(np.concatenate([att, rep]).astype(np.int64), device=dev)
    w = torch.tensor(np.concatenate([np.ones(len(att)), -np.ones(len(rep))]).astype(np.float32), device=dev)
    X = torch.randn(n, m, device=dev)
    res = {}
    for lay in ("soa", "ell"):
        os.environ["MDE_B200_LAYOUT"] = lay
        if rb:
            os.environ["MDE_B200_TILE_RB"] = rb
        else:
            os.environ.pop("MDE_B200_TILE_RB", None)
        for name, f in (("pp", pm.penalties.PushAndPull(w, pm.penalties.Log1p, pm.penalties.Log)),
                        ("hub", pm.losses.Huber(w.abs() * 0.7, 0.5))):
            mde = pm.MDE(n, m, edges, f, pm.Centered(), device=dev)
            Xg = X.clone().requires_grad_(True)
            v = mde.average_distortion(Xg)
            v.backward()
            torch.cuda.synchronize()
            res[(lay, name)] = (v.item(), Xg.grad.clone(), int(lib.mde_edges_kind(mde._layout().handle)))
    for name in ("pp", "hub"):
        a, b = res[("soa", name)], res[("ell", name)]
        print(n, m, name, "kinds", a[2], b[2], "value", a[0], b[0], "rel", abs(a[0] - b[0]) / abs(a[0]),
              "grad max diff / max", float((a[1] - b[1]).abs().max() / a[1].abs().max()), flush=True)
print("tiny done")
