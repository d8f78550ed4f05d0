#!/usr/bin/env python
"""One line: samples/s of the fused forward+adjoint sweep on a config-5 shard, with or without the explicit
persisting-L2 set-up (ADSURF_PERSIST_L2=0 skips adsurf_init).

    python benchmarks/quick_line.py [models] [steps]
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
d, a, b, rho, t, c = (torch.tensor(x, device="cuda") for x in config5(S, 128, 1024, 20))
out = tuple(torch.empty_like(d) for _ in range(4))
ws = torch.empty(_capi.lib().adsurf_workspace_bytes(_capi.OP_GRID_BWD, S, 20, 128, 1024), dtype=torch.uint8, device="cuda")
for _ in range(3):
    _capi.grid_bwd(d, a, b, rho, t, c, None, out=out, ws=ws)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    _capi.grid_bwd(d, a, b, rho, t, c, None, out=out, ws=ws)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"persist_l2={os.environ.get('ADSURF_PERSIST_L2', '1')} models={S}: {S * 128 * 1024 / ms / 1e6:.4f} G samples/s "
      f"({ms:.2f} ms per sweep, det not written)")
