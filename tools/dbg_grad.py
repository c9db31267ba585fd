import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
import pymde_b200 as pm
from oracle import mde_oracle as O
g = dict(np.load('tests/golden/trajectories.npz'))
key = 'pp_std'
w = g[key + '/par0']; e = g[key + '/edges']; X0 = g[key + '/X0']
f = pm.penalties.PushAndPull(torch.tensor(w, device='cuda'), pm.penalties.Log1p, pm.penalties.Log)
mde = pm.MDE(X0.shape[0], X0.shape[1], torch.tensor(e, device='cuda'), f, pm.Standardized())
X = torch.tensor(X0, device='cuda', requires_grad=True)
v = mde.average_distortion(X); v.backward()
spec = O.FnSpec(O.P_LOG1P, w, (1.5, 0, 0), fn_rep=O.P_LOG, rep=(1.0, 0, 0))
vr, gr = O.average_distortion(X0.astype(np.float64), e, spec, True)
err = np.abs(X.grad.cpu().numpy() - gr)
print(os.environ.get('TAG'), 'value', v.item(), vr, 'grad max err', err.max(), 'at', np.unravel_index(err.argmax(), err.shape), 'scale', np.abs(gr).max(), 'n bad rows', (err.max(1) > 1e-6).sum())
bad = np.where(err.max(1) > 1e-6)[0][:10]
print(' bad rows', bad.tolist())
