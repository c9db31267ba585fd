#!/usr/bin/env python
"""One-GPU runs of the other BASELINE configs (synthetic stand-ins): C3-shaped
(Scholar-like preserve_distances, Huber loss, Standardized) and a C4-shaped slice (m = 128, PushAndPull, Centered).
Prints kernel time, algorithmic GB/s, iterations/s and a parity spot check against the C oracle on a bounded
sample of edges."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pymde_b200 as pm
from pymde_b200 import _lib
from oracle import c_oracle, mde_oracle as O

dev = torch.device("cuda", 0)
lib = _lib.load()
peak = 6569.3


def kernel_ms(mde, X, reps=6):
    lay = mde._layout(); st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.zeros_like(X); ts = []
    for i in range(reps):
        g.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _lib.check(lib.mde_distortion(lay.handle, X.data_ptr(), X.shape[1], g.data_ptr(), None, st)); b.record()
        torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


def run(name, n, m, edges, f, cons, spec_fn, k=1, iters=20):
    mde = pm.MDE(n, m, edges, f, cons, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    X0 = torch.randn(n, m, device=dev, generator=gen)
    X0 = cons.project_onto_constraint(X0, inplace=True) if m <= 32 or cons is pm.Centered() else X0 - X0.mean(0)
    p = edges.shape[0]
    ms = kernel_ms(mde, X0)
    b_alg = p * (8 + 4 * k) + 2 * n * m * 4 + 8
    # parity spot check: value of a 200k-edge sample
    idx = torch.randperm(p, device=dev, generator=gen)[:200000].sort().values
    sub = pm.MDE(n, m, edges[idx], spec_fn(idx)[0], cons, device=dev)
    v = sub.average_distortion(X0).item()
    v_ref, _ = c_oracle.average_distortion(X0.cpu().numpy(), edges[idx].cpu().numpy(), spec_fn(idx)[1], want_grad=False)
    t0 = time.perf_counter(); mde.embed(X=X0, max_iter=iters, eps=0.0); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = mde.solve_stats
    print(json.dumps({"config": name, "n": n, "m": m, "p": p, "kernel_ms": ms, "algorithmic_gbs": b_alg / ms / 1e6,
                      "frac": b_alg / ms / 1e6 / peak, "iters_per_sec": st.iterations / dt,
                      "sample_value": v, "oracle_value": v_ref, "rel_err": abs(v - v_ref) / abs(v_ref),
                      "loss_first_last": [st.average_distortions[0], st.average_distortions[-1]]}))


# C3: n = 44 682, sampled pairs with small integer deviations rescaled to the standardized natural length
n, m, p = 44682, 2, 20_000_000
g = torch.Generator(device=dev); g.manual_seed(0)
e = torch.randint(0, n, (p, 2), device=dev, generator=g); e = e[e[:, 0] != e[:, 1]]
delta = torch.randint(1, 9, (e.shape[0],), device=dev, generator=g).float()
delta = pm.preprocess.scale(delta, pm.Standardized().natural_length(n, m))
run("C3-shaped: sampled pairs, losses.Huber(0.5), Standardized", n, m, e, pm.losses.Huber(delta, 0.5), pm.Standardized(),
    lambda idx: (pm.losses.Huber(delta[idx], 0.5), O.FnSpec(O.L_HUBER, delta[idx].cpu().numpy(), (0.5, 0, 0))))
del e, delta
# C4 slice: n = 1e6, m = 128, 15 pseudo-neighbours + as many repulsive pairs
n, m = 1_000_000, 128
i = torch.arange(n, device=dev).repeat_interleave(5)
j = (i + torch.randint(1, 1000, (i.numel(),), device=dev, generator=g)) % n
rep = torch.randint(0, n, (i.numel(), 2), device=dev, generator=g); rep = rep[rep[:, 0] != rep[:, 1]]
e = torch.cat([torch.stack([i, j], 1), rep]); w = torch.cat([torch.ones(i.numel(), device=dev), -torch.ones(rep.shape[0], device=dev)])
run("C4-shaped slice: m=128, PushAndPull(Log1p,Log), Centered", n, m, e,
    pm.penalties.PushAndPull(w, pm.penalties.Log1p, pm.penalties.Log), pm.Centered(),
    lambda idx: (pm.penalties.PushAndPull(w[idx], pm.penalties.Log1p, pm.penalties.Log),
                 O.FnSpec(O.P_LOG1P, w[idx].cpu().numpy(), (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0))), iters=10)
