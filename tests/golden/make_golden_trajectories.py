"""Generate tests/golden/trajectories.npz and profiles/r2_reference_configs_cpu.json by running the UNMODIFIED
reference's own `inversion(...)` (/root/reference, `_ADsurf.py:651`) here, on the CPU, for BASELINE configs 2-4.

    python tests/golden/make_golden_trajectories.py

SURVEY.md §8(d): "for configs 2-4 also report reference it/s on CPU for >= 5 iterations and compare loss
trajectories".  The reference is absent on the GPU box, so its trajectories travel as a fixture and the GPU test
(`tests/test_gpu_api.py::test_inversion_trajectories_match_the_reference`) replays the same inputs through
`adsurf_b200.inversion(...)`.  Two legs per config:
  * float64 (the reference's code with `numpy2tensor` rebound to float64 and the default dtype set to float64, the
    parity oracle of SURVEY.md §8c): losses and iterates of the first iterations -> the fixture;
  * native float32, exactly as shipped: it/s on this container's cores -> the timing record.
Config 4's fixture keeps the first 16 models of the 1024-model ensemble (models are independent without horizontal
damping, so each model's trajectory is the same in any ensemble); its it/s is timed on the full 1024 models.
Nothing in tests/, bench.py or the package imports this file.
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (stubs matplotlib, puts /root/reference on the path, fp64_reference)

import ADsurf._ADsurf as RA  # noqa: E402
import ADsurf._MonteCarlo as RMC  # noqa: E402

G._MODS.extend([RA, RMC])
NIT = 8


def mp6():
    return RA.Model_param(dc=0.001, vmin=2, vmax=3.2, tmin=1 / 10, tmax=5, layering_method="LN",
                          initialize_method="Brocher", layering_ratio=1.2, depth_factor=3.8, layer_number=6,
                          vp_vs_ratio=2.45, rho=2)


def run(tag, fn, its):
    t0 = time.perf_counter()
    res = fn(its)
    dt = time.perf_counter() - t0
    inv = res[0] if isinstance(res, tuple) else res
    loss = np.asarray(inv.inv_process["loss"], dtype=np.float64)
    print(f"{tag}: {its} iterations in {dt:.1f} s ({its / dt:.3f} it/s), loss[0] {np.ravel(loss[0])[:3]}")
    return res, its / dt, dt


def config2(its, f64):
    base = os.path.join(G.REF, "examples/data")
    obs = np.loadtxt(os.path.join(base, "00_increase_layer/input/disper.txt"))
    mp = mp6()
    ip = RA.inv_param(inversion_method="vs-and-thick", compress=True, compress_method="exp", normalized=True, lr=0.005,
                      damp_vertical=0.002, damp_horizontal=0, iteration=its, step_size=400, gamma=0.75)
    init = RA.Init_model(model_param=mp, pvs_obs=obs, thick=[0.4, 0.5, 0.5, 0.5, 0.5, 0.5])
    return RA.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=obs, vsrange_sign="mul",
                        vsrange=[0.1, 2], AK135_data=[], device="cpu"), init


def config3_models():
    base = os.path.join(G.REF, "examples/data")
    obs = np.loadtxt(os.path.join(base, "01_sandwich_layer/input/disper.txt"))
    mp = RA.Model_param(dc=0.001, vmin=0.2, vmax=0.4, tmin=1 / 100, tmax=1 / 5, layering_method="LN",
                        initialize_method="Constant", layering_ratio=2.5, depth_factor=2.5, layer_number=3,
                        vp_vs_ratio=2.45, rho=2, fundamental_range=[1 / 37, 1 / 8])
    init = RA.Init_model(model_param=mp, pvs_obs=obs, thick=np.array([0.005, 0.005, 0.005]))
    np.random.seed(123)
    mc = RA.init_model_MonteCarlo(mp, init, sampling_num=100, sampling_method="normal", vsrange_sign="plus",
                                  vsrange=[-0.1, 0.2], sigma_vs=20, sigma_thick=20, forward=False)
    M = {k: np.asarray(mc.MonteCarlo_model[k], dtype=np.float64) for k in ("vs", "vp", "thick", "rho")}
    return obs, mp, init, M


def config3(its, models):
    obs, mp, init, M = models
    ip = RA.inv_param(inversion_method="vs-and-thick", compress=True, compress_method="exp", normalized=True, lr=0.001,
                      damp_vertical=0.0, damp_horizontal=0.0, iteration=its, step_size=200, gamma=0.75)
    for key in ("vs", "vp", "thick", "rho"):
        init.init_model[key] = M[key].copy()
    S = M["vs"].shape[0]
    pvs = np.ones((S, obs.shape[0], 2)) * obs[:, :2]
    # torch.device object, not the string: the string takes the reference's broken cumsum branch
    return RA.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=pvs, vsrange_sign="plus",
                        vsrange=[-0.8, 0.8], AK135_data=[], device=torch.device("cpu"))


def config4(its, S):
    sys.path.insert(0, ROOT)
    from adsurf_b200.workload import config5_ensemble  # inputs only: bit-identical to the reference's generator
    base = os.path.join(G.REF, "examples/data")
    obs = np.loadtxt(os.path.join(base, "02_FRIUL7W/input/disper.txt"))
    d, a, b, rho = config5_ensemble(1024, 20, seed=0)
    d, a, b, rho = d[:S], a[:S], b[:S], rho[:S]
    mp = RA.Model_param(dc=0.01, vmin=0.5, vmax=5.0, tmin=1 / 20, tmax=5, layering_method="None",
                        initialize_method="Brocher", layering_ratio=1.2, depth_factor=2, layer_number=6,
                        vp_vs_ratio=2.45, rho=2)
    ip = RA.inv_param(inversion_method="vs", compress=False, compress_method="exp", normalized=True, lr=0.03,
                      damp_vertical=0.0, damp_horizontal=0.0, iteration=its, step_size=200, gamma=0.75)
    init = RA.Init_model(model_param=mp, pvs_obs=obs, vp=a[0], vs=b[0], rho=rho[0], thick=d[0])
    for key, val in (("vs", b), ("vp", a), ("thick", d), ("rho", rho)):
        init.init_model[key] = val.copy()
    pvs = np.ones((S, obs.shape[0], 2)) * obs[:, :2]
    inv = RA.inversion(model_param=mp, inv_param=ip, init_model=init, pvs_obs=pvs, vsrange_sign="mul",
                       vsrange=[0.1, 2.5], device=torch.device("cpu"))
    return inv, (d, a, b, rho)


def main():
    torch.set_num_threads(os.cpu_count())
    out, rec = {}, {"what": "the reference's own inversion(...) on this container's CPU (unmodified code), BASELINE "
                    "configs 2-4, first iterations", "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(),
                    "torch": torch.__version__}
    hist = lambda inv, key: np.asarray(inv.inv_process[key], dtype=np.float64)
    # ---- float64 legs -> fixture
    with G.fp64_reference():
        (inv, init), ips, _ = run("config2 f64", lambda n: config2(n, True), NIT)
        out.update(c2_loss=hist(inv, "loss"), c2_iter_vs=hist(inv, "iter_vs"), c2_iter_thick=hist(inv, "iter_thick"),
                   c2_init_vs=np.asarray(init.init_model["vs"], dtype=np.float64),
                   c2_init_thick=np.asarray(init.init_model["thick"], dtype=np.float64))
        rec["config2_float64"] = {"iterations": NIT, "it_per_s": ips}
        models = config3_models()
        inv, ips, _ = run("config3 f64", lambda n: config3(n, models), NIT)
        M = models[3]
        out.update(c3_loss=hist(inv, "loss"), c3_iter_vs=hist(inv, "iter_vs"), c3_iter_thick=hist(inv, "iter_thick"),
                   c3_vs=M["vs"], c3_vp=M["vp"], c3_thick=M["thick"], c3_rho=M["rho"])
        rec["config3_float64"] = {"iterations": NIT, "models": 100, "it_per_s": ips}
        (inv, (d, a, b, rho)), ips, _ = run("config4 f64 (16 models)", lambda n: config4(n, 16), NIT)
        out.update(c4_loss=hist(inv, "loss"), c4_iter_vs=hist(inv, "iter_vs"), c4_d=d, c4_a=a, c4_b=b, c4_rho=rho)
        rec["config4_float64_16_models"] = {"iterations": NIT, "models": 16, "it_per_s": ips}
    # ---- native float32 legs -> it/s
    (inv, _), ips, _ = run("config2 f32", lambda n: config2(n, False), 5)
    rec["config2_float32"] = {"iterations": 5, "it_per_s": ips, "first_losses": hist(inv, "loss")[:5].tolist()}
    models = config3_models()
    inv, ips, _ = run("config3 f32", lambda n: config3(n, models), 5)
    rec["config3_float32"] = {"iterations": 5, "models": 100, "it_per_s": ips,
                              "median_first_losses": np.median(hist(inv, "loss")[:5], axis=1).tolist()}
    (inv, _), ips, _ = run("config4 f32 (1024 models)", lambda n: config4(n, 1024), 5)
    rec["config4_float32"] = {"iterations": 5, "models": 1024, "it_per_s": ips,
                              "median_first_losses": np.median(hist(inv, "loss")[:5], axis=1).tolist()}
    np.savez_compressed(os.path.join(HERE, "trajectories"), **out)
    print({k: v.shape for k, v in out.items()})
    with open(os.path.join(ROOT, "profiles", "r2_reference_configs_cpu.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
