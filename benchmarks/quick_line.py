#!/usr/bin/env python
"""One line: samples/s of the fused forward+adjoint sweep on a config-5 shard, with or without the explicit
persisting-L2 set-up (ADSURF_PERSIST_L2=0 skips adsurf_init).

    python benchmarks/quick_line.py [models] [steps] [det]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from adsurf_b200 import _capi  # noqa: E402
from adsurf_b200.workload import config5  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
want_det = len(sys.argv) > 3 and sys.argv[3] == "det"
d, a, b, rho, t, c = (torch.tensor(x, device="cuda") for x in config5(S, 128, 1024, 20))
out = tuple(torch.empty_like(d) for _ in range(4))
ws = torch.empty(_capi.lib().adsurf_workspace_bytes(_capi.OP_GRID_BWD, S, 20, 128, 1024), dtype=torch.uint8, device="cuda")
import ctypes  # noqa: E402

det = torch.empty((S, 1024, 128), dtype=torch.float64, device="cuda") if want_det else None
P = lambda x: None if x is None else ctypes.c_void_p(x.data_ptr())
_capi.ensure_init(d.device)


def sweep():
    rc = _capi.lib().adsurf_det_grid_bwd(P(d), P(a), P(b), P(rho), P(t), P(c), None, S, 20, 128, 1024, P(det), P(out[0]),
                                         P(out[1]), P(out[2]), P(out[3]), P(ws), ws.numel(),
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc


for _ in range(3):
    sweep()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    sweep()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"persist_l2={os.environ.get('ADSURF_PERSIST_L2', '1')} models={S}: {S * 128 * 1024 / ms / 1e6:.4f} G samples/s "
      f"({ms:.2f} ms per sweep, det {'written' if want_det else 'not written'})")
