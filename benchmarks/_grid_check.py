"""dev helper: dense fwd+adjoint of the current ADSURF_GRID_BWD mode against the default on a small case"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adsurf_b200 import _capi
from adsurf_b200.workload import config5
d, a, b, rho, t, c = config5(5, 100, 70)
G = lambda x: torch.tensor(x, device="cuda")
out = _capi.grid_bwd(G(d), G(a), G(b), G(rho), G(t), G(c), None, want_det=True)
np.save(sys.argv[1], np.concatenate([o.cpu().numpy().ravel() for o in out]))
