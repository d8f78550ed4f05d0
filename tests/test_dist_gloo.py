"""N > 1 path on the CPU: world_size-2 gloo, ensemble sharded over the model axis.

The sharded `multi_inversion` (rows split across ranks, horizontal-damping halo, final all-gather)
must reproduce the single-process run exactly — both the device-resident loop (`InversionLoop`, which exchanges
four Vs rows per iteration) and the torch-autograd loop; the two loops must agree with each other.  C-ABI calls are served by the oracle (fake backend);
this tests the host-side sharding logic, not the kernels."""
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import load


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _MP:  # minimal monkeypatch stand-in usable inside spawned workers
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _problem(S=5, damp_horizontal=0.003, iteration=3):
    g = load("dmf_exp_m6")
    rng = np.random.default_rng(3)
    b = np.tile(g["b"][0], (S, 1)) * (1 + rng.normal(0, 0.02, (S, 6)))
    d = np.tile(g["d"][0], (S, 1))
    from tests.golden_models import brocher_np
    a, rho = brocher_np(b)
    obs = np.stack([np.column_stack((g["t"][0][:12], g["c"][0][:12]))] * S)
    mp_ = types.SimpleNamespace(clist=np.arange(2, 3.2, 0.05), layering_method="LN", initialize_method="Brocher",
                                vp_vs_ratio=2.45, rho=2)
    ip = types.SimpleNamespace(damp_vertical=0.002, damp_horizontal=damp_horizontal, compress=True, normalized=True,
                               compress_method="exp", inversion_method="vs-and-thick", lr=0.01, iteration=iteration,
                               step_size=2, gamma=0.5, optimizer="Adam")
    init = types.SimpleNamespace(init_model={"vs": b, "vp": a, "rho": rho, "thick": d,
                                             "layer_mindepth": np.full(6, 0.05), "layer_maxdepth": np.full(6, 4.0)})
    return mp_, ip, init, obs


def _run(fused=True, **problem):
    from adsurf_b200._model import multi_inversion
    mp_, ip, init, obs = _problem(**problem)
    os.environ["ADSURF_FUSED_OPT"] = "1" if fused else "0"
    inv = multi_inversion(mp_, ip, init, obs, vsrange_sign="add", vsrange=[-0.5, 0.5], device="cpu")
    assert inv.timing["fused"] == fused
    return (np.asarray(inv.inv_process["loss"]), np.asarray(inv.inv_process["iter_vs"]),
            np.asarray(inv.inv_process["iter_thick"]), np.asarray(inv.inv_model["vs"]))


def _worker(rank, world, port, out_dir):
    torch.set_num_threads(1)
    from tests import fake_backend
    fake_backend.install(_MP())
    ref = _run()  # single-process result (device-resident loop), before the process group exists
    ref_torch = _run(fused=False)  # torch-autograd loop
    for x, y in zip(ref, ref_torch):
        np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-12)
    even = dict(S=4, damp_horizontal=0.0, iteration=5)
    ref_even = _run(**even)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from adsurf_b200 import _dist
        assert _dist.is_sharded() and _dist.shard_bounds(5) == ((0, 3) if rank == 0 else (3, 5))
        got = _run()  # sharded, device-resident loop with the four-row halo exchange (horizontal damping > 0)
        for x, y in zip(got, ref):
            np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-14)
        got_torch = _run(fused=False)  # sharded torch-autograd loop (full all-gather of Vs per iteration)
        for x, y in zip(got_torch, ref_torch):
            np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-14)
        # equal shards without horizontal damping (graph-replay form of the loop, one collective per history array)
        got_even = _run(**even)
        for x, y in zip(got_even, ref_even):
            np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-14)
        # uneven all-gather helper
        lo, hi = _dist.shard_bounds(5)
        full = _dist.all_gather_rows(torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1), 5)
        assert torch.equal(full.reshape(-1), torch.arange(5, dtype=torch.float64))
        # equal shards take the single-collective path; gathered along a middle axis
        lo, hi = _dist.shard_bounds(4)
        part = torch.arange(lo, hi, dtype=torch.float64).reshape(1, -1, 1).expand(3, -1, 2).contiguous()
        full = _dist.all_gather_rows(part, 4, dim=1)
        assert full.shape == (3, 4, 2) and torch.equal(full[1, :, 0], torch.arange(4, dtype=torch.float64))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_inversion_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_bounds_cover_everything():
    from adsurf_b200._dist import shard_bounds
    for S in (1, 2, 7, 100, 4096):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(S, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == S
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
