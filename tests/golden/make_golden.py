"""Generate golden fixtures by RUNNING THE UNMODIFIED REFERENCE (cvxgrp/pymde v0.2.1).

Run in the build container (where /root/reference or baseline/_ref exists):

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/functions.npz, evals.npz, projections.npz,
trajectories.npz.  The GPU box has no reference; tests there read these files.
Every array is produced by reference code paths only:
  functions    f(d) and autograd df/dd   pymde/functions/{penalties,losses}.py
  evals        MDE.average_distortion + backward   pymde/average_distortion.py:36-80
  projections  constraints + util.proj_standardized  pymde/constraints.py, pymde/util.py:129-171
  trajectories MDE.embed                 pymde/problem.py:386 -> pymde/optim.py:69 -> pymde/lbfgs.py
"""
import functools
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle.ref_loader import load_reference  # noqa: E402

pymde = load_reference()
assert pymde is not None, "reference not importable"
torch.set_num_threads(1)  # bit-reproducible summation order


def function_cases(p, dtype, rng):
    """name -> (constructor kwargs as arrays, callable building the reference module)."""
    pen, los = pymde.penalties, pymde.losses
    w = torch.tensor(rng.uniform(0.2, 2.0, p), dtype=dtype)
    wneg = -w
    wmix = torch.tensor(rng.choice([1.0, 2.0, -1.0], p), dtype=dtype)
    dev = torch.tensor(rng.uniform(0.1, 3.0, p), dtype=dtype)
    w2 = torch.tensor(rng.uniform(0.2, 2.0, p), dtype=dtype)
    return {
        "pen_linear": (pen.Linear(w), dict(par0=w)),
        "pen_quadratic": (pen.Quadratic(w), dict(par0=w)),
        "pen_cubic": (pen.Cubic(w), dict(par0=w)),
        "pen_power_2.5": (pen.Power(w, 2.5), dict(par0=w)),
        "pen_huber_0.5": (pen.Huber(w, 0.5), dict(par0=w)),
        "pen_logistic_0.3_3": (pen.Logistic(w, 0.3, 3.0), dict(par0=w)),
        "pen_log1p_1.5": (pen.Log1p(w, 1.5), dict(par0=w)),
        "pen_log_1": (pen.Log(wneg, 1.0), dict(par0=wneg)),
        "pen_invpower_1": (pen.InvPower(wneg, 1), dict(par0=wneg)),
        "pen_logratio_2": (pen.LogRatio(wneg, 2), dict(par0=wneg)),
        "pen_pushpull_log1p_log": (pen.PushAndPull(wmix, pen.Log1p, pen.Log), dict(par0=wmix)),
        "pen_pushpull_default": (pen.PushAndPull(wmix), dict(par0=wmix)),
        "pen_pushpull_quad_invpower": (
            pen.PushAndPull(wmix, pen.Quadratic, pen.InvPower), dict(par0=wmix)),
        "loss_absolute": (los.Absolute(dev), dict(par0=dev)),
        "loss_quadratic": (los.Quadratic(dev), dict(par0=dev)),
        "loss_weighted_quadratic": (los.WeightedQuadratic(dev), dict(par0=dev)),
        "loss_weighted_quadratic_w": (los.WeightedQuadratic(dev, w2), dict(par0=dev, par1=w2)),
        "loss_huber_0.7": (los.Huber(dev, 0.7), dict(par0=dev)),
        "loss_cubic": (los.Cubic(dev), dict(par0=dev)),
        "loss_power_1.5": (los.Power(dev, 1.5), dict(par0=dev)),
        "loss_logistic": (los.Logistic(dev), dict(par0=dev)),
        "loss_fractional": (los.Fractional(dev), dict(par0=dev)),
        "loss_soft_fractional_10": (los.SoftFractional(dev, 10.0), dict(par0=dev)),
    }


def gen_functions():
    out = {}
    p = 64
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        rng = np.random.default_rng(7)
        cases = function_cases(p, dtype, rng)
        d = torch.tensor(rng.uniform(0.05, 4.0, p), dtype=dtype)
        for name, (f, pars) in cases.items():
            dd = d.clone().requires_grad_(True)
            val = f(dd)
            val.sum().backward()
            out["%s/%s/d" % (name, tag)] = d.numpy()
            out["%s/%s/f" % (name, tag)] = val.detach().numpy()
            out["%s/%s/fp" % (name, tag)] = dd.grad.numpy()
            for k, v in pars.items():
                out["%s/%s/%s" % (name, tag, k)] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "functions.npz"), **out)
    print("functions.npz", len(out))


def random_graph(n, p, rng):
    pairs = set()
    while len(pairs) < p:
        i, j = rng.integers(0, n, 2)
        if i != j:
            pairs.add((min(i, j), max(i, j)))
    e = np.array(sorted(pairs), dtype=np.int64)
    rng.shuffle(e)
    return e


def gen_evals():
    out = {}
    rng = np.random.default_rng(11)
    for m, zero in ((1, False), (2, False), (3, False), (4, False), (7, False), (16, False),
                    (2, True), (3, True)):
        n, p = 60, 64
        edges = random_graph(n, p, rng)
        # flip some edges so i > j also occurs (tolerated, problem.py:80-83)
        flip = rng.random(p) < 0.3
        edges[flip] = edges[flip][:, ::-1]
        X = rng.standard_normal((n, m)).astype(np.float32)
        key = "m%d" % m
        if zero:  # a zero-distance pair is legal input (average_distortion.py:55-62)
            edges = edges[~((edges == 5).any(1) & (edges == 9).any(1))]
            edges[0] = (5, 9)
            X[5] = X[9]
            key = "m%d_zero" % m
        out[key + "/edges"] = edges
        out[key + "/X"] = X
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            cases = function_cases(p, dtype, np.random.default_rng(7))
            for name, (f, pars) in cases.items():
                mde = pymde.MDE(n, m, torch.tensor(edges), f, pymde.Centered())
                Xt = torch.tensor(X, dtype=dtype, requires_grad=True)
                v = mde.average_distortion(Xt)
                v.backward()
                out["%s/%s/%s/value" % (key, name, tag)] = v.detach().numpy()
                out["%s/%s/%s/grad" % (key, name, tag)] = Xt.grad.numpy()
                if m == 2 and tag == "f32" and not zero:
                    out["%s/%s/distances" % (key, name)] = mde.distances(Xt.detach()).numpy()
                    out["%s/%s/distortions" % (key, name)] = mde.distortions(Xt.detach()).numpy()
    np.savez_compressed(os.path.join(HERE, "evals.npz"), **out)
    print("evals.npz", len(out))


def gen_projections():
    out = {}
    rng = np.random.default_rng(3)
    std, cen = pymde.Standardized(), pymde.Centered()
    for (n, m) in ((2, 2), (10, 3), (100, 3), (1000, 2), (257, 5), (300, 40)):
        Z = rng.standard_normal((n, m)).astype(np.float32) + 0.3
        key = "n%d_m%d" % (n, m)
        out[key + "/Z"] = Z
        out[key + "/centered"] = cen.project_onto_constraint(torch.tensor(Z), inplace=False).numpy()
        Xs = std.project_onto_constraint(torch.tensor(Z).clone(), inplace=False)
        out[key + "/standardized"] = Xs.numpy()
        G = rng.standard_normal((n, m)).astype(np.float32)
        out[key + "/G"] = G
        out[key + "/tangent"] = std.project_onto_tangent_space(
            Xs, torch.tensor(G), inplace=False).numpy()
    np.savez_compressed(os.path.join(HERE, "projections.npz"), **out)
    print("projections.npz", len(out))


def knn_like_graph(n, k, rng):
    """Synthetic stand-in for a k-NN graph on a ring with a few chords (no datasets here)."""
    e = set()
    for i in range(n):
        for o in range(1, k + 1):
            j = (i + o) % n
            e.add((min(i, j), max(i, j)))
    for _ in range(n // 2):
        i, j = rng.integers(0, n, 2)
        if i != j:
            e.add((min(i, j), max(i, j)))
    return np.array(sorted(e), dtype=np.int64)


def gen_trajectories():
    out = {}
    pen, los = pymde.penalties, pymde.losses

    def run(key, n, m, edges, f, constraint, max_iter, X0=None, eps=1e-5, memory_size=10):
        mde = pymde.MDE(n, m, torch.tensor(edges), f, constraint)
        if X0 is None:
            torch.manual_seed(0)
            X0 = constraint.initialization(n, m)
        # float64 run of the same reference code (tensors stay float64 end to end): pins the
        # ALGORITHM (L-BFGS, Wolfe search, projections) free of fp32 summation-order noise
        f64 = f.double()
        mde64 = pymde.MDE(n, m, torch.tensor(edges), f64, constraint)
        X64 = mde64.embed(X=X0.double(), max_iter=max_iter, eps=eps, memory_size=memory_size)
        out[key + "/f64/X"] = X64.detach().numpy()
        out[key + "/f64/average_distortions"] = np.array(mde64.solve_stats.average_distortions)
        out[key + "/f64/residual_norms"] = np.array(mde64.solve_stats.residual_norms)
        out[key + "/f64/step_size_percents"] = np.array(mde64.solve_stats.step_size_percents)
        f.float()
        X = mde.embed(X=X0, max_iter=max_iter, eps=eps, memory_size=memory_size)
        st = mde.solve_stats
        out[key + "/edges"] = edges
        out[key + "/X0"] = X0.numpy()
        out[key + "/X"] = X.detach().numpy()
        out[key + "/average_distortions"] = np.array(st.average_distortions)
        out[key + "/residual_norms"] = np.array(st.residual_norms)
        out[key + "/step_size_percents"] = np.array(st.step_size_percents)
        out[key + "/final_value"] = mde.average_distortion(X.detach()).numpy()
        out[key + "/max_iter"] = np.array(max_iter)
        out[key + "/eps"] = np.array(eps)
        print(key, "iters", st.iterations, "value", float(out[key + "/final_value"]))

    rng = np.random.default_rng(5)
    # T1: quadratic + standardized (spectral problem; converges)
    n, m = 80, 2
    edges = knn_like_graph(n, 3, rng)
    w = np.ones(len(edges), np.float32)
    out["quad_std/par0"] = w
    run("quad_std", n, m, edges, pen.Quadratic(torch.tensor(w)), pymde.Standardized(), 60)
    # T2: push-and-pull Log1p/Log + centered (the preserve_neighbors default)
    n, m = 120, 2
    att = knn_like_graph(n, 4, rng)
    attset = set(map(tuple, att))
    rep = []
    while len(rep) < len(att):
        i, j = rng.integers(0, n, 2)
        if i != j and (min(i, j), max(i, j)) not in attset:
            rep.append((min(i, j), max(i, j)))
            attset.add((min(i, j), max(i, j)))
    edges = np.concatenate([att, np.array(rep, dtype=np.int64)])
    w = np.concatenate([rng.choice([1.0, 2.0], len(att)), -np.ones(len(rep))]).astype(np.float32)
    out["pp_cen/par0"] = w
    run("pp_cen", n, m, edges, pen.PushAndPull(torch.tensor(w), pen.Log1p, pen.Log),
        pymde.Centered(), 40)
    # T3: push-and-pull + standardized, m = 3
    out["pp_std/par0"] = w
    run("pp_std", n, 3, edges, pen.PushAndPull(torch.tensor(w), pen.Log1p, pen.Log),
        pymde.Standardized(), 40)
    # T4: C1-shaped: cycle graph, all pairs, hop distances, Absolute, centered (nonsmooth)
    n, m = 40, 2
    iu = np.triu_indices(n, 1)
    edges = np.stack(iu, 1).astype(np.int64)
    hop = np.minimum(edges[:, 1] - edges[:, 0], n - (edges[:, 1] - edges[:, 0])).astype(np.float32)
    out["cycle_abs/par0"] = hop
    run("cycle_abs", n, m, edges, los.Absolute(torch.tensor(hop)), pymde.Centered(), 40)
    # T5: C3-shaped: Huber loss + standardized on sampled pairs
    n, m = 100, 2
    edges = random_graph(n, 900, rng)
    delta = rng.integers(1, 6, len(edges)).astype(np.float32)
    delta = delta * (np.sqrt(2.0 * n * m / (n - 1)) / np.sqrt((delta ** 2).mean()))
    delta = delta.astype(np.float32)
    out["huber_std/par0"] = delta
    run("huber_std", n, m, edges, los.Huber(torch.tensor(delta), 0.5), pymde.Standardized(), 40)
    # T6: docs example (docs_src/source/mde/index.rst:193-286): 5 items, 4 edges
    edges = np.array([[0, 1], [0, 2], [0, 3], [3, 4]], dtype=np.int64)
    w = np.array([1.0, 2.0, 5.0, 6.0], np.float32)
    out["docs5/par0"] = w
    run("docs5", 5, 2, edges, pen.Quadratic(torch.tensor(w)), pymde.Standardized(), 100)
    np.savez_compressed(os.path.join(HERE, "trajectories.npz"), **out)
    print("trajectories.npz", len(out))


if __name__ == "__main__":
    gen_functions()
    gen_evals()
    gen_projections()
    gen_trajectories()
