#!/usr/bin/env python
"""CPU arm of BASELINE configs 2-4 on THIS machine's host cores: the same `inversion(...)` calls as
benchmarks/configs_e2e.py with the C ABI served by the float64 oracle port (tests/fake_backend.py ->
oracle/dltar_oracle.py: torch-CPU restatement of the reference's evaluators, autograd gradients, the package's own
host loop).  SURVEY.md §8(d) asks for the reference's it/s on the box's CPU over >= 5 iterations with the loss
trajectories compared: the reference itself cannot travel to the GPU box (its own runs, timed in the build container,
are profiles/r2_reference_configs_cpu.json and the fixture tests/golden/trajectories.npz), so the port — which
reproduces the reference's trajectories to 4e-12 (tests/test_host_logic.py) — is timed here.  bench.py runs this file
in a subprocess (N = 1 only) and compares the first losses with its GPU runs of the same configs.

    python benchmarks/cpu_port_configs.py [--iters 5] [--models4 1024]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

import torch  # noqa: E402


class _MP:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--models4", type=int, default=1024)
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    from tests import fake_backend
    fake_backend.install(_MP())  # every C-ABI entry point served by the oracle on the CPU
    import configs_e2e as CE
    out = {"kind": "port", "cores": os.cpu_count(), "threads": torch.get_num_threads(), "iterations": args.iters}
    for name, fn in (("config2", lambda: CE.config2(args.iters)), ("config3", lambda: CE.config3(args.iters)),
                     ("config4", lambda: CE.config4(args.iters, S=args.models4, arrays_too=False))):
        r = fn()
        secs = r.get("loop_seconds", r.get("seconds"))
        out[name] = {"it_per_s": args.iters / secs, "seconds": secs, "first_losses": r["first_losses"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
