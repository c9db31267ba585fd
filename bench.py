#!/usr/bin/env python
"""bench.py -- headline benchmark of the MDE hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                    # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path, host cores

A "step" is one iteration of MDE.embed(): >= 1 fused evaluation of the whole edge list (value + gradient),
the projection(s), the L-BFGS update and the strong-Wolfe line search.

N = 1   BASELINE.json configs[1] (C2): MNIST-shaped preserve_neighbors, n = 70 000, m = 2, ~1.55 M edges,
        PushAndPull(Log1p(1.5), Log(1.0)), Centered -- synthetic (no datasets / network here).
N > 1   BASELINE.json configs[4] (C5): synthetic SBM, n = 10 000 000, m = 2, PushAndPull, Centered, ONE problem
        cut into edge shards of 25 000 000 edges; rank r owns shard r (weak scaling: p = 2.5e7 N, n fixed), X is
        replicated and the (n, m) gradient is all-reduced once per evaluation by the library's own
        peer-memory kernels over NVLink.  The line also carries the same shard solved on ONE GPU
        (`single_gpu`: the weak-scaling base), C2 cut into N shards (`c2_sharded`: the latency cost of the
        exchange on a 560 KB gradient) and a `parity` object.

JSON keys beyond the base contract: `roofline` (fused distortion kernel, cold L2, CUDA events), `cpu_baseline`
(the unmodified reference on the host cores, bounded sample), `e2e` (public API with pinned HOST buffers, copies
inside the timed region), `gpu_launches`, `clocks`, `iters_per_sec`, `reference_torch_cuda`,
`parity_at_equal_iterations`.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_ITEMS, EMBED_DIM, K_NEIGHBORS = 70000, 2, 15
C5_N, C5_SHARD_EDGES, C5_BLOCK = 10_000_000, 25_000_000, 10_000
REPEATS = 5          # timed windows of K steps each; the line reports the median window
E2E_CALLS = 15       # wall-clock end-to-end calls (host-side noise -- e.g. the nvidia-smi sampler taking driver locks -- has
                     # a heavy tail at 4 ms per call: the median of 15 is stable where the median of 5 was not)
CPU_THREADS = 16     # ATen's CPU scatter_add / index kernels stop scaling near 16 threads (r01: 16 beat 32, 128)


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
def c2_edges(seed, n=N_ITEMS, k=K_NEIGHBORS):
    """Attractive: per item k pseudo-neighbours with index locality inside 10 'classes',
    symmetrised + de-duplicated the way Graph.from_edges does (duplicates summed => w in {1,2},
    sorted by (i,j); pymde/preprocess/graph.py:21-72).  Repulsive: as many uniformly sampled
    non-neighbour pairs, w = -1, appended unsorted (pymde/recipes.py:388-416)."""
    rng = np.random.default_rng(seed)
    blob = n // 10
    i = np.repeat(np.arange(n, dtype=np.int64), k)
    # two-sided local offsets: ~half of the directed picks are reciprocated, as in a real k-NN
    # graph (MNIST: 70 000 x 15 directed -> 776 k undirected edges, examples/mnist.ipynb:56)
    off = rng.integers(1, 24, n * k) * rng.choice([-1, 1], n * k)
    j = (i // blob) * blob + ((i % blob) + off) % blob
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    key = lo * n + hi
    uniq, counts = np.unique(key, return_counts=True)
    att = np.stack([uniq // n, uniq % n], 1)
    w_att = counts.astype(np.float32).clip(max=2.0)
    n_rep = len(att)
    cand = rng.integers(0, n, (int(n_rep * 1.2), 2))
    cand = cand[cand[:, 0] != cand[:, 1]]
    ck = np.minimum(cand[:, 0], cand[:, 1]) * n + np.maximum(cand[:, 0], cand[:, 1])
    ck = ck[~np.isin(ck, uniq)]
    _, first = np.unique(ck, return_index=True)
    ck = ck[np.sort(first)][:n_rep]
    rep = np.stack([ck // n, ck % n], 1)
    edges = np.concatenate([att, rep]).astype(np.int64)
    w = np.concatenate([w_att, -np.ones(len(rep), np.float32)])
    return edges, w


def initial_iterate(seed, n=N_ITEMS, m=EMBED_DIM):
    rng = np.random.default_rng(seed)
    X0 = rng.standard_normal((n, m)).astype(np.float32)
    return X0 - X0.mean(0)


def c5_shard(shard, n=C5_N, p=C5_SHARD_EDGES, block=C5_BLOCK):
    """Shard `shard` of the C5 problem: stochastic block model over n nodes in blocks of 10 000;
    half of the shard's edges attractive (w = +1; 90 % inside the block of their first endpoint, 10 % anywhere),
    half repulsive uniform pairs (w = -1).  Seeded by the shard index only, so the problem does not depend on
    the number of GPUs: an N-GPU run solves the union of shards 0..N-1."""
    rng = np.random.default_rng([5, shard])
    pa = p // 2
    i = rng.integers(0, n, pa, dtype=np.int64)
    j_in = (i // block) * block + rng.integers(0, block, pa, dtype=np.int64)
    j_out = rng.integers(0, n, pa, dtype=np.int64)
    j = np.where(rng.random(pa) < 0.9, j_in, j_out)
    np.minimum(j, n - 1, out=j)
    keep = i != j
    att = np.stack([i[keep], j[keep]], 1)
    del i, j, j_in, j_out, keep
    rep = rng.integers(0, n, (p - pa, 2), dtype=np.int64)
    rep = rep[rep[:, 0] != rep[:, 1]]
    edges = np.concatenate([att, rep])
    w = np.concatenate([np.ones(len(att), np.float32), -np.ones(len(rep), np.float32)])
    return edges, w


def make_config(world):
    if world == 1:
        return {"workload": "C2: MNIST-shaped preserve_neighbors (synthetic): n=%d, m=%d, ~%d neighbours, "
                            "PushAndPull(Log1p(1.5), Log(1.0)), Centered" % (N_ITEMS, EMBED_DIM, K_NEIGHBORS),
                "n_items": N_ITEMS, "embedding_dim": EMBED_DIM, "memory_size": 10,
                "parallelism": "1 GPU",
                "l2": "solver loop runs L2-warm (working set ~26 MB < 126 MB L2); roofline kernel timed cold (512 MB flush)"}
    return {"workload": "C5: synthetic SBM, n=%d, m=2, %d edges per shard x %d shards (one shard per GPU), "
                        "PushAndPull(Log1p(1.5), Log(1.0)), Centered" % (C5_N, C5_SHARD_EDGES, world),
            "n_items": C5_N, "embedding_dim": 2, "memory_size": 10,
            "parallelism": "edge-sharded x%d, X replicated, 1 peer-memory all-reduce of the (n,m) gradient per evaluation" % world,
            "l2": "inputs larger than L2 (300 MB of edge records + 160 MB of X and gradient per GPU): no flush needed"}


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows for k in range(4) if len(r) > 2 + k and r[2 + k] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------
# reference arms (the UNMODIFIED reference from baseline/_ref; the numpy oracle port only if it did not travel)
# ------------------------------------------------------------------------------------------
def cpu_reference_run(n, m, edges, w, X0, iters, warm):
    """`warm` untimed + `iters` timed embed iterations of the reference on the host cores, `CPU_THREADS` torch
    threads.  Returns dict(value=edges/s, iters_per_sec, cores, kind, seconds)."""
    import torch
    from oracle.ref_loader import load_reference
    cores = os.cpu_count() or 1
    ref = load_reference()
    p = len(edges)
    if ref is not None:
        threads = min(cores, CPU_THREADS)
        torch.set_num_threads(threads)
        f = ref.penalties.PushAndPull(torch.tensor(w), ref.penalties.Log1p, ref.penalties.Log)
        mde = ref.MDE(n, m, torch.tensor(edges), f, ref.Centered(), device="cpu")
        if warm:
            mde.embed(X=torch.tensor(X0), max_iter=warm, eps=0.0)
        t0 = time.perf_counter()
        mde.embed(X=torch.tensor(X0), max_iter=iters, eps=0.0)
        dt = time.perf_counter() - t0
        done, kind = mde.solve_stats.iterations, "reference"
    else:
        from oracle import mde_oracle as O
        spec = O.FnSpec(O.P_LOG1P, w, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0))
        t0 = time.perf_counter()
        _, st = O.embed(X0, edges, spec, O.Centered(), eps=0.0, max_iter=iters, dtype=np.float32)
        dt = time.perf_counter() - t0
        done, kind, threads = st.iterations, "port", 1
    ips = done / dt
    return {"value": ips * p, "unit": "edges/s", "iters_per_sec": ips, "cores": threads, "kind": kind,
            "host_cores": cores, "seconds": dt, "iterations": done}


def cuda_reference_run(edges, w, X0, iters, dev):
    """The reference's own torch-CUDA path (device='cuda') on the same GPU: the denominator of the
    north_star's '10x the reference's torch-CUDA embed() steps/sec'.  None when baseline/_ref is absent."""
    import torch
    from oracle.ref_loader import load_reference
    ref = load_reference()
    if ref is None:
        return None
    try:
        f = ref.penalties.PushAndPull(torch.tensor(w, device=dev), ref.penalties.Log1p, ref.penalties.Log)
        mde = ref.MDE(X0.shape[0], X0.shape[1], torch.tensor(edges, device=dev), f, ref.Centered(), device=dev)
        X0d = torch.tensor(X0, device=dev)
        for _ in range(2):
            mde.embed(X=X0d, max_iter=5, eps=0.0)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        mde.embed(X=X0d, max_iter=iters, eps=0.0)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        done = mde.solve_stats.iterations
        return {"iters_per_sec": done / dt, "value": done / dt * len(edges), "unit": "edges/s", "iterations": done,
                "average_distortions": [float(v) for v in mde.solve_stats.average_distortions],
                "what": "unmodified reference, device='cuda', same edges/weights/X0, eps=0"}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)}


def reference_arm(args, rank, world, K, W):
    """--impl reference: the reference's CPU implementation of the path on the host cores, same metric / config /
    steps / warmup as our arm.  N = 1: the full C2 workload.  N > 1: a bounded sample of the C5 workload (the first
    tenth of shard 0 over all 10M nodes) so that K + W iterations end within minutes."""
    if rank != 0:
        return 0
    if world == 1:
        edges, w = c2_edges(0)
        X0 = initial_iterate(0)
        n, m = N_ITEMS, EMBED_DIM
        sample = "the full C2 workload (n=%d, p=%d), %d warm-up + %d timed embed iterations from the same X0" % (
            n, len(edges), W, K)
    else:
        edges, w = c5_shard(0)
        cut = len(edges) // 10
        sel = np.concatenate([np.arange(cut // 2), len(edges) - 1 - np.arange(cut // 2)])  # both classes
        edges, w = edges[sel], w[sel]
        n, m = C5_N, 2
        X0 = initial_iterate(2, n, m)
        sample = ("bounded sample of the C5 workload: %d edges (a tenth of shard 0, both classes) over all n=%d nodes, "
                  "%d warm-up + %d timed embed iterations" % (len(edges), n, W, K))
    r = cpu_reference_run(n, m, edges, w, X0, K, W)
    line = {"impl": "reference", "metric": "embed_edges_per_sec", "value": r["value"], "unit": "edges/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 / r["iters_per_sec"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "iters_per_sec": r["iters_per_sec"], "config": make_config(world),
            "cpu_baseline": {"value": r["value"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": sample + "; %d torch threads on a %d-core host" % (r["cores"], r["host_cores"])},
            "e2e": {"value": r["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def timed_windows(solver, K, repeats, barrier, torch, dev, world):
    """`repeats` windows of exactly K solver iterations, each bracketed by barrier + synchronize, CUDA events
    on the launching stream; the max over ranks of every window.  Returns (list of ms, iterations done)."""
    out = []
    done = 0
    for _ in range(repeats):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        done, _ = solver.run(K)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            import torch.distributed as tdist
            t = torch.tensor([ms], device=dev)
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = float(t.item())
        out.append(ms)
    return out, done


def load_profile(key):
    """dram traffic of the dominant kernel from the committed ncu artifact (profiles/r02_kernel_profile.json,
    written by tools/ncu_extract.py from the --set full capture); None when the artifact has no such entry."""
    try:
        prof = json.load(open(os.path.join(REPO, "profiles", "r02_kernel_profile.json")))
        return prof.get(key)
    except Exception:
        return None


def c3_shaped_roofline(torch, dev, lib, _lib, pm):
    """The fused kernel on the shape of BASELINE configs[2] (~40k nodes, dense preserve_distances pairs, losses.Huber):
    n = 44 682, 2e7 random pairs, m = 2 -- the dense case the library gives the ELL pull records (kind 3).  Same timing
    as `roofline` (CUDA events around one launch, 512 MB L2 flush before each; the 240 MB record stream is larger than
    L2 anyway).  A second roofline object, not the headline: bench's step stays C2."""
    n, m, p = 44682, 2, 20_000_000
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    e = torch.randint(0, n, (p, 2), device=dev, generator=gen)
    e = e[e[:, 0] != e[:, 1]]
    delta = torch.randint(1, 9, (e.shape[0],), device=dev, generator=gen).float() * 0.25
    t0 = time.perf_counter()
    mde = pm.MDE(n, m, e, pm.losses.Huber(delta, 0.5), pm.Centered(), device=dev)
    X = torch.randn(n, m, device=dev, generator=gen)
    X -= X.mean(0)
    mde._layout()
    torch.cuda.synchronize(dev)
    build_s = time.perf_counter() - t0
    r = kernel_roofline(mde, X, m, int(e.shape[0]), n, torch, dev, lib, _lib, True, "C3_shaped", reps=10)
    kind = int(lib.mde_edges_kind(mde._layout().handle))
    r["kernel"] = {0: "distortion_quad_kernel<m=2, fused, L_HUBER> (sorted-SoA layout)",
                   2: "distortion_pull_kernel<m=2, fused, L_HUBER> (pull-record layout)",
                   3: "distortion_ell_kernel<m=2, fused, L_HUBER> (ELL pull records, one lane per owner)"}.get(kind, "?")
    r["workload"] = "C3-shaped: n=44682, m=2, %d random pairs, losses.Huber(threshold 0.5)" % int(e.shape[0])
    r["layout_build_seconds"] = build_s
    r["timing"] = "CUDA events around one launch, 512 MB L2 flush before each, mean of 10"
    del mde
    torch.cuda.empty_cache()
    return r


def kernel_roofline(mde, X0d, m, p_local, n, torch, dev, lib, _lib, cold, profile_key, reps=20):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    lay = mde._layout()
    grad = torch.zeros_like(X0d)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if cold else None  # > 126 MB L2
    times = []
    st = torch.cuda.current_stream(dev).cuda_stream
    for it in range((reps + 4) if cold else 10):
        if cold:
            flush.fill_(it & 0xFF)
        grad.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.mde_distortion(lay.handle, X0d.data_ptr(), m, grad.data_ptr(), None, st))
        b.record()
        torch.cuda.synchronize(dev)
        if it >= 4:
            times.append(a.elapsed_time(b))
    k_ms = float(np.mean(times))
    b_alg = p_local * 12 + 2 * n * m * 4 + 8
    achieved = b_alg / (k_ms * 1e-3) / 1e9
    kind = int(lib.mde_edges_kind(lay.handle))  # the layout / kernel family the library picked for this workload
    prof = load_profile("%s:%s" % (profile_key, {0: "soa", 1: "tiles", 2: "pull", 3: "ell"}.get(kind, "?")))
    kernel = {0: "distortion_quad_kernel<m=2, fused, LOG1P|LOG, fast-math> (sorted-SoA layout)",
              1: "distortion_tile_kernel<m=2, fused, LOG1P|LOG, fast-math> (tile-record layout, push)",
              2: "distortion_pull_kernel<m=2, fused, LOG1P|LOG, fast-math> (pull-record layout)",
              3: "distortion_ell_kernel<m=2, fused, LOG1P|LOG, fast-math> (ELL pull records, one lane per owner)"
              }.get(kind, "?")
    return {"bound": "hbm", "kernel": kernel,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None if prof is None else prof.get("dram_bytes"),
            "traffic_source": None if prof is None else prof.get("source"),
            "ncu_duration_us": None if prof is None else prof.get("ncu_duration_us"),  # same capture, cold caches
            "peak_source": peak_src, "algorithmic_bytes": b_alg, "kernel_ms": k_ms,
            "timing": ("CUDA events around one launch, 512 MB L2 flush before each, mean of 20" if cold else
                       "CUDA events around one launch, inputs larger than L2, mean of 6")}


def tensor_digest(t):
    return hashlib.sha1(t.detach().cpu().numpy().tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-iters", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3)

    if args.impl == "reference":
        return reference_arm(args, rank, world, K, W)

    import torch
    import pymde_b200 as pm
    from pymde_b200 import _lib, dist as pdist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    tdist = None
    if world > 1:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize(dev)

    if world == 1:
        return ours_single(args, K, W, torch, pm, _lib, lib, dev, barrier)
    return ours_sharded(args, K, W, rank, world, torch, tdist, pm, pdist, _lib, lib, dev, barrier)


def ours_single(args, K, W, torch, pm, _lib, lib, dev, barrier):
    edges, w = c2_edges(0)
    p = len(edges)
    X0 = initial_iterate(0)
    wt = torch.tensor(w, device=dev)
    f = pm.penalties.PushAndPull(wt, pm.penalties.Log1p, pm.penalties.Log)
    mde = pm.MDE(N_ITEMS, EMBED_DIM, torch.tensor(edges, device=dev), f, pm.Centered(), device=dev)
    X0d = torch.tensor(X0, device=dev)

    # -- device-resident throughput: W warm-up, then REPEATS windows of K timed iterations ----------------
    solver = mde._solver(mde.constraint, 10, REPEATS * K + W + 8)
    sampler = ClockSampler(dev.index or 0)
    solver.begin(X0d, 0.0)
    solver.run(W)
    barrier()
    sampler.start()
    launches0 = lib.mde_launch_count()
    windows, done = timed_windows(solver, K, REPEATS, barrier, torch, dev, 1)
    launches = (lib.mde_launch_count() - launches0) / REPEATS
    ms = float(np.median(windows))
    avg, res, pct, stp, fe = solver.stats(done)

    # -- end to end through the public API with pinned HOST buffers ---------------------------------------
    X0h = torch.tensor(X0).pin_memory()
    out_h = torch.empty_like(X0h).pin_memory()
    mde.embed(X=X0h, max_iter=W, eps=0.0)  # warm the API path
    e2e_runs = []
    for _ in range(E2E_CALLS):
        barrier()
        t0 = time.perf_counter()
        Xe = mde.embed(X=X0h, max_iter=K, eps=0.0)
        out_h.copy_(Xe, non_blocking=True)
        torch.cuda.synchronize(dev)
        e2e_runs.append(time.perf_counter() - t0)
    e2e_s = float(np.median(e2e_runs))
    e2e_iters = mde.solve_stats.iterations
    ours_traj = [float(v) for v in mde.solve_stats.average_distortions]
    h2d = X0h.numel() * 4
    d2h = out_h.numel() * 4 + 4 * 8 * e2e_iters + 48 * (e2e_iters // 64 + 2)  # X, statistics, status words

    roofline = kernel_roofline(mde, X0d, EMBED_DIM, p, N_ITEMS, torch, dev, lib, _lib, True, "C2")
    roofline_c3 = None
    try:  # an extra measurement must never cost the bench line
        roofline_c3 = c3_shaped_roofline(torch, dev, lib, _lib, pm)
    except Exception as ex:
        roofline_c3 = {"error": repr(ex)[:200]}
    clocks = sampler.stop()  # sampled across all timed regions above (solver windows, e2e calls, kernel timing)

    cpu_baseline = None
    if not args.no_cpu_baseline:
        r = cpu_reference_run(N_ITEMS, EMBED_DIM, edges, w, X0, args.cpu_iters, 2)
        cpu_baseline = {"value": r["value"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
                        "iters_per_sec": r["iters_per_sec"],
                        "sample": "%d embed iterations of the full workload (n=%d, p=%d) from the same X0 after 2 warm-up "
                                  "iterations; %d torch threads on a %d-core host" % (
                                      r["iterations"], N_ITEMS, p, r["cores"], r["host_cores"])}

    ref_cuda, parity_eq = None, None
    if not args.no_cpu_baseline:
        ref_cuda = cuda_reference_run(edges, w, X0, K, dev)
        if ref_cuda and "average_distortions" in ref_cuda:
            rt = ref_cuda.pop("average_distortions")
            k_cmp = min(len(rt), len(ours_traj))
            parity_eq = {"iterations": k_cmp,
                         "what": "average distortion logged at the start of iteration i, same X0 / edges / weights, both fp32 "
                                 "on this GPU; non-converged trajectories of an fp32 quasi-Newton method drift apart "
                                 "(the reference's own run-to-run spread with different thread counts is 2e-4 .. 4e-3, "
                                 "SURVEY App. C.4); the 1e-5 criterion is tested on converged problems in tests/",
                         "ours": {str(i): ours_traj[i] for i in (0, 1, 2, 5, 10, k_cmp - 1) if i < k_cmp},
                         "reference_torch_cuda": {str(i): rt[i] for i in (0, 1, 2, 5, 10, k_cmp - 1) if i < k_cmp},
                         "rel_diff_first": abs(ours_traj[0] - rt[0]) / abs(rt[0]),
                         "rel_diff_last": abs(ours_traj[k_cmp - 1] - rt[k_cmp - 1]) / abs(rt[k_cmp - 1])}

    ips = K / (ms * 1e-3)
    cfg = make_config(1)  # identical in both arms (the driver compares it); run details go to top-level keys
    line = {
        "metric": "embed_edges_per_sec", "value": ips * p, "unit": "edges/s", "n_gpus": 1, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "iters_per_sec": ips, "timed_windows_ms": windows, "func_evals_total": int(fe),
        "edges_total": p, "edges_per_gpu": p, "repeats": REPEATS,
        "average_distortion_after_%d_iterations" % done: float(avg[-1]) if len(avg) else None,
        "e2e": {"value": e2e_iters / e2e_s * p, "unit": "edges/s", "iters_per_sec": e2e_iters / e2e_s,
                "h2d_bytes_per_step": h2d / max(e2e_iters, 1), "d2h_bytes_per_step": d2h / max(e2e_iters, 1),
                "seconds_per_call": e2e_runs,
                "what": "MDE.embed(X=pinned host X0, max_iter=K) + copy of the embedding to pinned host memory; "
                        "median of %d calls" % E2E_CALLS},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_c3": roofline_c3,
        "cpu_baseline": cpu_baseline,
        "reference_torch_cuda": ref_cuda, "parity_at_equal_iterations": parity_eq,
    }
    print(json.dumps(line))
    return 0


def ours_sharded(args, K, W, rank, world, torch, tdist, pm, pdist, _lib, lib, dev, barrier):
    n, m = C5_N, 2
    edges, w = c5_shard(rank)
    p_local = len(edges)
    t = torch.tensor([p_local], device=dev, dtype=torch.int64)
    tdist.all_reduce(t)
    p_total = int(t.item())
    ed = torch.tensor(edges, device=dev)
    wt = torch.tensor(w, device=dev)
    del edges, w
    X0 = initial_iterate(2, n, m)
    X0d = torch.tensor(X0, device=dev)

    def build(shard_of_global):
        f = pm.penalties.PushAndPull(wt, pm.penalties.Log1p, pm.penalties.Log)
        mde = pm.MDE(n, m, ed, f, pm.Centered(), device=dev)
        if shard_of_global:
            pdist.attach(mde, rank, world, p_total, dev)
        return mde

    # -- weak-scaling base: the same shard solved alone on this GPU (no exchange) --------------------------
    Kb = min(K, 20)
    solo = build(False)
    s1 = solo._solver(solo.constraint, 10, Kb + W + 8)
    s1.begin(X0d, 0.0)
    s1.run(W)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.run(Kb)
    e1.record()
    torch.cuda.synchronize(dev)
    solo_ms = e0.elapsed_time(e1) / Kb
    roofline = kernel_roofline(solo, X0d, m, p_local, n, torch, dev, lib, _lib, False, "C5_shard") if rank == 0 else None
    tt = torch.tensor([solo_ms], device=dev)
    tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
    solo_ms = float(tt.item())
    s1.close()
    solo.__dict__["_device_solver"] = None
    del s1

    # -- the sharded solve ----------------------------------------------------------------------------------
    mde = build(True)
    solver = mde._solver(mde.constraint, 10, REPEATS * K + W + 8)
    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    solver.begin(X0d, 0.0)
    solver.run(W)
    barrier()
    if sampler:
        sampler.start()
    launches0 = lib.mde_launch_count()
    windows, done = timed_windows(solver, K, REPEATS, barrier, torch, dev, world)
    launches = (lib.mde_launch_count() - launches0) / REPEATS
    clocks = sampler.stop() if sampler else None
    ms = float(np.median(windows))
    avg, res, pct, stp, fe = solver.stats(done)

    # -- parity: the library's peer-memory all-reduce against NCCL on the same per-shard partials; replicas agree ---
    lay = mde._layout()
    g = torch.zeros_like(X0d)
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    _lib.check(lib.mde_distortion(lay.handle, X0d.data_ptr(), m, g.data_ptr(), loss.data_ptr(),
                                  torch.cuda.current_stream(dev).cuda_stream))
    g64 = g.double()
    tdist.all_reduce(g64)
    tdist.all_reduce(loss)
    loss0_nccl = float(loss.item()) / p_total
    resid0_nccl = float(g64.norm().item())
    digest = tensor_digest(solver.x_view())
    digests = [None] * world
    tdist.all_gather_object(digests, digest)
    del g, g64

    # -- end to end: pinned host X0 in, embedding out --------------------------------------------------------
    X0h = torch.tensor(X0).pin_memory()
    out_h = torch.empty_like(X0h).pin_memory()
    Ke = min(K, 20)
    mde.embed(X=X0h, max_iter=3, eps=0.0)
    barrier()
    t0 = time.perf_counter()
    Xe = mde.embed(X=X0h, max_iter=Ke, eps=0.0)
    out_h.copy_(Xe, non_blocking=True)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    tt = torch.tensor([e2e_s], device=dev)
    tdist.all_reduce(tt, op=tdist.ReduceOp.MAX)
    e2e_s = float(tt.item())
    e2e_iters = mde.solve_stats.iterations
    h2d = X0h.numel() * 4
    d2h = out_h.numel() * 4 + 4 * 8 * e2e_iters + 48 * (e2e_iters // 64 + 2)
    peer = bool(getattr(solver, "peer_memory", False))
    del mde, solver, solo
    torch.cuda.empty_cache()

    # -- C2 cut into `world` shards: what the exchange costs on a latency-bound 560 KB gradient -------------
    c2 = None
    try:
        e2, w2 = c2_edges(0)
        lo, hi = pdist.shard_range(len(e2), rank, world)
        w2t = torch.tensor(w2[lo:hi], device=dev)
        m2 = pm.MDE(N_ITEMS, 2, torch.tensor(e2[lo:hi], device=dev),
                    pm.penalties.PushAndPull(w2t, pm.penalties.Log1p, pm.penalties.Log), pm.Centered(), device=dev)
        pdist.attach(m2, rank, world, len(e2), dev)
        s2 = m2._solver(m2.constraint, 10, 3 * 100 + 16)
        s2.begin(torch.tensor(initial_iterate(0), device=dev), 0.0)
        s2.run(5)
        w2ms, _ = timed_windows(s2, 100, 3, barrier, torch, dev, world)
        c2 = {"workload": "C2 (1.55 M edges) cut into %d edge shards" % world, "ms_per_step": float(np.median(w2ms)) / 100,
              "iters_per_sec": 100 / (float(np.median(w2ms)) * 1e-3)}
    except Exception as ex:  # pragma: no cover
        c2 = {"error": repr(ex)[:200]}

    if rank != 0:
        return 0
    ips = K / (ms * 1e-3)
    cfg = make_config(world)  # identical in both arms (the driver compares it); run details go to top-level keys
    line = {
        "metric": "embed_edges_per_sec", "value": ips * p_total, "unit": "edges/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "iters_per_sec": ips, "timed_windows_ms": windows, "func_evals_total": int(fe),
        "edges_total": p_total, "edges_per_gpu": p_local, "repeats": REPEATS,
        "allreduce": "peer-memory kernels (cudaIpc + flag handshake, graph-captured)" if peer else "NCCL host hook",
        "single_gpu": {"what": "rank-local shard (%d edges, n=%d) solved alone on one GPU, max over ranks" % (p_local, n),
                       "ms_per_step": solo_ms, "value": p_local / (solo_ms * 1e-3), "unit": "edges/s"},
        "c2_sharded": c2,
        "parity": {"x_bit_identical_across_ranks": len(set(digests)) == 1,
                   "loss_at_x0": float(avg[0]), "loss_at_x0_nccl_fp64": loss0_nccl,
                   "loss_rel_diff": abs(float(avg[0]) - loss0_nccl) / abs(loss0_nccl),
                   "grad_norm_at_x0": float(res[0]), "grad_norm_at_x0_nccl_fp64": resid0_nccl,
                   "grad_norm_rel_diff": abs(float(res[0]) - resid0_nccl) / abs(resid0_nccl),
                   "what": "iteration-0 loss and ||gradient|| of the sharded solver (peer-memory all-reduce, fp32) against an "
                           "NCCL fp64 all-reduce of the same per-shard kernel outputs; the C-oracle comparison of the sharded "
                           "evaluation lives in tests/test_gpu_multi.py"},
        "e2e": {"value": e2e_iters / e2e_s * p_total, "unit": "edges/s", "iters_per_sec": e2e_iters / e2e_s,
                "h2d_bytes_per_step": h2d / max(e2e_iters, 1), "d2h_bytes_per_step": d2h / max(e2e_iters, 1),
                "what": "MDE.embed(X=pinned host X0, max_iter=%d) on every rank + copy of the embedding to pinned host memory" % Ke},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": None,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
