#!/usr/bin/env python
"""it/s of the device-resident ensemble loop (config-4 shape: 300 picks, 20 layers, Vs only) against the number of
concurrent row groups (ADSURF_LOOP_GROUPS, adsurf_b200._capi.InversionLoopGroup) for several ensemble sizes.

    python benchmarks/loop_groups.py [--sizes 512,1024,4096] [--groups 0,1,2,3,4]   (0 = the package's own choice) [--iters 300]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="512,1024,4096")
    ap.add_argument("--groups", default="1,2,3,4")
    ap.add_argument("--iters", type=int, default=300)
    args = ap.parse_args()
    import configs_e2e as CE
    CE.config4(5, S=64)
    ref = {}
    for S in (int(x) for x in args.sizes.split(",")):
        for G in (int(x) for x in args.groups.split(",")):
            os.environ["ADSURF_LOOP_GROUPS"] = str(G) if G > 0 else ""   # 0: the package's own choice
            CE.config4(5, S=S)
            best = None
            for _ in range(2):
                r = CE.config4(args.iters, S=S)
                if best is None or r["it_per_s"] > best["it_per_s"]:
                    best = r
            key = (S,)
            ref.setdefault(key, best["median_last_loss"])
            print(json.dumps({"models": S, "groups": G, "it_per_s": round(best["it_per_s"], 1),
                              "us_per_iteration": round(1e6 / best["it_per_s"], 1),
                              "median_last_loss": best["median_last_loss"],
                              "same_result_as_first": bool(np.isclose(best["median_last_loss"], ref[key], rtol=1e-9))}),
                  flush=True)


if __name__ == "__main__":
    main()
