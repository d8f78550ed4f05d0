#!/usr/bin/env python
"""SURVEY.md §8(d) CPU baseline as specified: the REFERENCE's own code, unmodified, timed in the build container.

    python benchmarks/time_reference_cpu.py > profiles/r2_reference_cpu_timing.json

Runs /root/reference's `_cps/_surf96_matrixAll_gpu.dltar_matrix(..., device="cpu")` + `.backward()` on chunks of the
config-5 workload (8 models x 128 f x kc c x 20 layers; it cannot hold the full grid: ~33 kB of autograd state per
sample) in native float32 and in float64 (numpy2tensor rebound + default dtype float64, exactly as
tests/golden/make_golden.py does), with all host threads, and the oracle port (oracle/dltar_oracle.py — what
bench.py times on the GPU box, where /root/reference does not exist) on the same chunk for comparison.
Also times the reference's multi_inversion-style step for configs 2-4?  No: those loops are dominated by the same
call; their notebook rates are in BASELINE.md.  Only runnable where /root/reference is mounted.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import make_golden as MG  # noqa: E402  (stubs matplotlib, imports the reference, provides fp64_reference)
from adsurf_b200.workload import config5_ensemble, config5_periods, config5_velocities  # noqa: E402
from oracle import dltar_oracle as O  # noqa: E402


def chunk(models, kc):
    d, a, b, rho = config5_ensemble(max(models, 2), 20, 0)
    d, a, b, rho = (x[:models].copy() for x in (d, a, b, rho))
    t = np.tile(config5_periods(128), (models, 1))
    c = np.tile(config5_velocities(1024)[:: 1024 // kc][:kc], (models, 1))
    return d, a, b, rho, t, c


def ref_fwd_bwd(args, dtype, backward=True):
    d, a, b, rho, t, c = args
    leaves = [torch.tensor(x, dtype=dtype, requires_grad=backward) for x in (d, a, b, rho)]
    det = MG.ref_matall.dltar_matrix(torch.tensor(c, dtype=dtype), torch.tensor(t, dtype=dtype), *leaves, 2, -1,
                                     device="cpu")
    if backward:
        det.backward(torch.ones_like(det))
    return det.numel(), det.dtype


def port_fwd_bwd(args, backward=True):
    d, a, b, rho, t, c = args
    leaves = [torch.tensor(x, requires_grad=backward) for x in (d, a, b, rho)]
    det = O.det_grid(torch.tensor(c), torch.tensor(t), *leaves)
    if backward:
        det.backward(torch.ones_like(det))
    return det.numel(), det.dtype


def timed(fn, budget=20.0, max_n=20):
    fn()  # warm-up
    n, t0, samples = 0, time.perf_counter(), 0
    while True:
        k, dt = fn()
        samples += k
        n += 1
        el = time.perf_counter() - t0
        if el > budget or n >= max_n:
            return {"samples_per_s": samples / el, "repeats": n, "seconds": el, "dtype": str(dt)}


def main():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    models, kc = 8, 64   # 8 x 128 x 64 = 65 536 samples per call (the float64 reference needs ~2.3 GB for it)
    args = chunk(models, kc)
    out = {"what": "reference CPU baseline of SURVEY.md §8(d): /root/reference ADsurf/_cps/_surf96_matrixAll_gpu.py "
                   "dltar_matrix(device='cpu') + backward(ones), unmodified code, build container",
           "cores": cores, "torch_threads": torch.get_num_threads(), "torch": torch.__version__,
           "chunk": f"{models} models x 128 f x {kc} c x 20 layers of config 5 ({models * 128 * kc} samples per call)"}
    out["reference_float32_fwd_bwd"] = timed(lambda: ref_fwd_bwd(args, torch.float32))
    out["reference_float32_fwd_only"] = timed(lambda: ref_fwd_bwd(args, torch.float32, backward=False))
    with MG.fp64_reference():
        out["reference_float64_fwd_bwd"] = timed(lambda: ref_fwd_bwd(args, torch.float64))
        out["reference_float64_fwd_only"] = timed(lambda: ref_fwd_bwd(args, torch.float64, backward=False))
    out["oracle_port_float64_fwd_bwd"] = timed(lambda: port_fwd_bwd(args))
    out["oracle_port_float64_fwd_only"] = timed(lambda: port_fwd_bwd(args, backward=False))
    out["port_over_reference_float64"] = (out["oracle_port_float64_fwd_bwd"]["samples_per_s"]
                                          / out["reference_float64_fwd_bwd"]["samples_per_s"])
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
