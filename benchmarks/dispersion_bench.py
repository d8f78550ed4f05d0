#!/usr/bin/env python
"""Timing of the GPU dispersion-curve extraction: the reference's root search run by `k_dispersion`
(one warp per model, the lanes sharing each evaluation) for one model and for ensembles, next to the sequential host solver (`_surf96.surf96`, the
Python port with the reference's control flow) on this box.

    python benchmarks/dispersion_bench.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from adsurf_b200._surf96 import surf96  # noqa: E402
from adsurf_b200.dispersion import dispersion_curves  # noqa: E402
from adsurf_b200.workload import config5_ensemble, config5_periods  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / reps


def main():
    d, a, b, rho = config5_ensemble(4096, 20, seed=0)
    t = config5_periods(128)
    if len(sys.argv) > 2 and sys.argv[1] == "--profile":  # two launches of S models (for ncu)
        S = int(sys.argv[2])
        for _ in range(2):
            dispersion_curves(t, d[:S], a[:S], b[:S], rho[:S], modes=(0,), dc=0.01)
        torch.cuda.synchronize()
        return
    rec = {"what": "20-layer FRIUL7W ensemble of config 5, 128 periods, fundamental mode, dc = 0.01"}
    t0 = time.perf_counter()
    host = surf96(t, d[0], a[0], b[0], rho[0], 0, 0, 2, 0.01)
    rec["host_solver_1_model_s"] = time.perf_counter() - t0
    for S in (1, 64, 4096):
        c, dt = timed(lambda: dispersion_curves(t, d[:S] if S > 1 else d[0], a[:S] if S > 1 else a[0],
                                                b[:S] if S > 1 else b[0], rho[:S] if S > 1 else rho[0], modes=(0,), dc=0.01))
        rec[f"gpu_{S}_models_s"] = dt
        rec[f"gpu_{S}_models_curves_per_s"] = S / dt
        if S == 1:
            rec["max_abs_diff_vs_host_km_s"] = float(np.max(np.abs(c[0] - host)))
    c3, dt3 = timed(lambda: dispersion_curves(t, d[:4096], a[:4096], b[:4096], rho[:4096], modes=(0, 1, 2), dc=0.01), reps=1)
    rec["gpu_4096_models_3_modes_s"] = dt3
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
