#!/usr/bin/env python
"""Sharded vs unsharded `multi_inversion` on real GPUs (torchrun, NCCL): the device-resident loop with horizontal
damping (four-row halo exchange per iteration) and without it (CUDA-graph replay, concurrent row groups), each compared with the same
ensemble inverted whole on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        benchmarks/check_sharded.py
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def problem(S, damp_h):
    from adsurf_b200.workload import config5_ensemble
    obs = np.load(os.path.join(ROOT, "tests", "golden", "shipped_fixtures.npz"))["disper_02"][::3]
    d, a, b, rho = config5_ensemble(S, 20, seed=0)
    mp = types.SimpleNamespace(clist=np.arange(0.5, 5.0, 0.02), layering_method="None", initialize_method="Brocher",
                               vp_vs_ratio=2.45, rho=2)
    ip = types.SimpleNamespace(damp_vertical=0.002, damp_horizontal=damp_h, compress=True, normalized=True,
                               compress_method="exp", inversion_method="vs", lr=0.01, iteration=20, step_size=8,
                               gamma=0.75, optimizer="Adam")
    init = types.SimpleNamespace(init_model={"vs": b, "vp": a, "rho": rho, "thick": d,
                                             "layer_mindepth": np.full(20, 0.01), "layer_maxdepth": np.full(20, 50.0)})
    pvs = np.ones((S, obs.shape[0], 2)) * obs[:, :2]
    return mp, ip, init, pvs


def run(S, damp_h):
    from adsurf_b200._model import multi_inversion
    mp, ip, init, pvs = problem(S, damp_h)
    inv = multi_inversion(mp, ip, init, pvs, vsrange_sign="mul", vsrange=[0.5, 1.5], device="cuda")
    return (np.asarray(inv.inv_process["loss"]), np.asarray(inv.inv_process["iter_vs"]), np.asarray(inv.inv_model["vs"]),
            inv.timing)


def main():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"world": world, "backend": dist.get_backend()}
    # uneven shards (padded list all-gather) and even ones (single all_gather_into_tensor), with / without damping
    for name, S, dh in (("131 models, damp_h=0 (graph replay)", 64 * world + 3, 0.0),
                        ("131 models, damp_h=0.003 (halo exchange)", 64 * world + 3, 0.003),
                        ("128 models, damp_h=0.003 (halo exchange, even shards)", 64 * world, 0.003),
                        ("512 models, damp_h=0 (even shards, two concurrent row groups per rank)",
                         256 * world, 0.0)):
        got = run(S, dh)
        assert got[3]["fused"]
        dist.barrier()
        if rank == 0:
            os.environ["ADSURF_NO_SHARD"] = "1"
            ref = run(S, dh)
            os.environ["ADSURF_NO_SHARD"] = "0"
            err = [float(np.max(np.abs(x - y) / np.maximum(np.abs(y), 1e-300))) for x, y in zip(got[:3], ref[:3])]
            out[name] = {"max_rel_diff_loss": err[0], "max_rel_diff_iter_vs": err[1], "max_rel_diff_final_vs": err[2],
                         "loss_first_median": float(np.median(ref[0][0])), "loss_last_median": float(np.median(ref[0][-1]))}
            assert max(err) <= 1e-9, err
        dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
