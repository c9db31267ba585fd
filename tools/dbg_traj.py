import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
import pymde_b200 as pm
from tests.test_gpu_solver import build
g = dict(np.load('tests/golden/trajectories.npz'))
key = sys.argv[1]
mde, X0 = build(pm, key, g)
mde.embed(X=X0, max_iter=4, eps=1e-5)
st = mde.solve_stats
print(os.environ.get('TAG', ''), key, ['%.7f' % v for v in st.average_distortions], ['%.6g' % v for v in st.step_lengths], st.func_evals,
      'ref', ['%.7f' % v for v in g[key + '/average_distortions'][:4]])
