#!/usr/bin/env python
"""cProfile of one config-4 `inversion(...)` call (1024 models, 400 iterations): where the host-side time around the
device-resident loop goes.   python benchmarks/profile_config4.py [arrays]"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

import configs_e2e as C  # noqa: E402

if __name__ == "__main__":
    C.config4(40, 1024)  # warm-up: library load, graph capture paths, allocator
    pr = cProfile.Profile()
    pr.enable()
    C.config4(400, 1024)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
