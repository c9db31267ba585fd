"""Quick timing of other solver configurations on the C2-sized problem (not the headline bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, pymde_b200 as pm
dev = torch.device("cuda", 0)
edges, w = bench.c2_edges(0); X0 = bench.initial_iterate(0)
et = torch.tensor(edges, device=dev)
cases = {
  "pp_centered_m2": (pm.penalties.PushAndPull(torch.tensor(w, device=dev), pm.penalties.Log1p, pm.penalties.Log), pm.Centered(), 2),
  "pp_standardized_m2": (pm.penalties.PushAndPull(torch.tensor(w, device=dev), pm.penalties.Log1p, pm.penalties.Log), pm.Standardized(), 2),
  "pp_centered_m3": (pm.penalties.PushAndPull(torch.tensor(w, device=dev), pm.penalties.Log1p, pm.penalties.Log), pm.Centered(), 3),
  "huberloss_standardized_m2": (pm.losses.Huber(torch.tensor(np.abs(w) + 0.5, device=dev), 0.5), pm.Standardized(), 2),
  "quadratic_standardized_m2": (pm.penalties.Quadratic(torch.tensor(np.abs(w), device=dev)), pm.Standardized(), 2),
  "pp_centered_m16": (pm.penalties.PushAndPull(torch.tensor(w, device=dev), pm.penalties.Log1p, pm.penalties.Log), pm.Centered(), 16),
}
for name, (f, cons, m) in cases.items():
    mde = pm.MDE(bench.N_ITEMS, m, et, f, cons, device=dev)
    X0m = cons.initialization(bench.N_ITEMS, m, dev)
    solver = mde._solver(cons, 10, 400); solver.begin(X0m, 0.0); solver.run(5); torch.cuda.synchronize()
    t0 = time.perf_counter(); done, _ = solver.run(200); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    avg, res, pct, stp, fe = solver.stats(done)
    print("%-28s %8.0f it/s  evals/iter %.2f  loss %.5f -> %.5f" % (name, (done - 5) / dt, fe / done, avg[0], avg[-1]))
