"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) here.

Run once in the build container (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

Nothing in tests/, bench.py or the package imports this file.  What it does:
  * stubs matplotlib/seaborn so ``ADsurf`` imports;
  * float64 legs: rebinds ``numpy2tensor`` in ``ADsurf._utils`` and in every module that
    imported it by name to a float64 version and sets the default dtype to float64, so the
    reference's own code runs in double precision;
  * float32 legs: the reference exactly as shipped;
  * calls the reference's own ``dltar_vector`` / ``dltar_matrix`` / ``multi_cal_grad`` /
    ``iter_cal_grad`` / ``surf96`` and stores inputs, outputs and autograd gradients;
  * copies the handful of numbers needed from the reference's shipped fixtures
    (examples/data/**/disper*.txt rows, model.json loss/iterate heads).
"""
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "seaborn"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
sys.path.insert(0, REF)

import ADsurf._utils as U  # noqa: E402
import ADsurf._cps._surf96_vector as ref_vec  # noqa: E402
import ADsurf._cps._surf96_matrix as ref_mat  # noqa: E402
import ADsurf._cps._surf96_vector_gpu as ref_vec_gpu  # noqa: E402
import ADsurf._cps._surf96_matrix_gpu as ref_mat_gpu  # noqa: E402
import ADsurf._cps._surf96_vectorAll_gpu as ref_vecall  # noqa: E402
import ADsurf._cps._surf96_matrixAll_gpu as ref_matall  # noqa: E402
import ADsurf._model as ref_model  # noqa: E402
import ADsurf._surf96 as ref_surf  # noqa: E402

_MODS = [U, ref_vec, ref_mat, ref_vec_gpu, ref_mat_gpu, ref_vecall, ref_matall, ref_model]
_ORIG = U.numpy2tensor


def _n2t64(a):
    if not torch.is_tensor(a):
        return torch.tensor(a).to(torch.float64)
    return a.to(torch.float64)


class fp64_reference:
    def __enter__(self):
        for m in _MODS:
            if hasattr(m, "numpy2tensor"):
                m.numpy2tensor = _n2t64
        torch.set_default_dtype(torch.float64)

    def __exit__(self, *a):
        for m in _MODS:
            if hasattr(m, "numpy2tensor"):
                m.numpy2tensor = _ORIG
        torch.set_default_dtype(torch.float32)


def brocher(vs):
    vs = np.asarray(vs, dtype=np.float64)
    _, vp, vs, rho = U.gen_model(np.ones_like(vs), vs, area=True)
    return vp, rho


def log_wavelength_periods(tmin, tmax, n):
    return np.exp(np.linspace(np.log(tmin), np.log(tmax), n))


def t64(x, grad=False):
    t = torch.tensor(np.asarray(x, dtype=np.float64))
    t.requires_grad_(grad)
    return t


MODEL6 = dict(thick=[0.4, 0.5, 0.5, 0.5, 0.5, 0.5], vs=[2.39, 2.54, 2.7, 2.82, 2.93, 3.04])
MODEL3 = dict(thick=[0.005, 0.005, 0.005], vs=[0.3, 0.22, 0.4])


def friul(nl):
    tab = np.loadtxt(os.path.join(REF, "examples/data/02_FRIUL7W/input/FRIUL7W-structure.txt"))
    return dict(thick=tab[:nl, 1].tolist(), vs=tab[:nl, 5].tolist())


def ensemble(model, S, sigma, rng):
    vs = np.asarray(model["vs"])
    th = np.asarray(model["thick"])
    b = vs[None, :] + rng.normal(0, sigma, (S, len(vs)))
    b[0] = vs
    d = th[None, :] * (1 + rng.normal(0, 0.05, (S, len(vs))))
    d[0] = th
    a, rho = brocher(b)
    return d, a, b, rho


def case_grid_single(name, model, vlist, tlist):
    """single-model dltar_matrix [K,NF]: float64-patched and native float32 (CPU twin + _gpu twin)."""
    vp, rho = brocher(model["vs"])
    d, b = np.asarray(model["thick"], float), np.asarray(model["vs"], float)
    with fp64_reference():
        # the `_gpu` twin is used for float64: the CPU twin hard-casts vlist/omega to float32
        # (_cps/_surf96_matrix.py:260-261) so it cannot be lifted to double by rebinding numpy2tensor
        det64 = ref_mat_gpu.dltar_matrix(t64(vlist), t64(tlist), t64(d), t64(vp), t64(b), t64(rho), 2, -1, device="cpu")
    f32 = lambda x: torch.tensor(np.asarray(x, dtype=np.float32))
    det32 = ref_mat.dltar_matrix(f32(vlist), f32(tlist), f32(d), f32(vp), f32(b), f32(rho), 2, -1)
    assert det64.dtype == torch.float64 and det32.dtype == torch.float32
    np.savez_compressed(os.path.join(OUT, name), vlist=vlist, tlist=tlist, d=d, a=vp, b=b, rho=rho,
                        det64=det64.numpy(), det32=det32.numpy())
    print(name, det64.shape, float(det64.abs().max()))


def case_points_batched(name, model, S, N, crange, trange, seed, sigma=0.05, single_too=True):
    rng = np.random.default_rng(seed)
    d, a, b, rho = ensemble(model, S, sigma, rng)
    t = np.exp(rng.uniform(np.log(trange[0]), np.log(trange[1]), (S, N)))
    c = rng.uniform(crange[0], crange[1], (S, N))
    gF = rng.standard_normal((S, N))
    with fp64_reference():
        td, ta, tb, trho, tc = (t64(x, True) for x in (d, a, b, rho, c))
        F = ref_vecall.dltar_vector(tc, t64(t), td, ta, tb, trho, 2, -1, device="cpu")
        gc, gd, ga, gb, grho = torch.autograd.grad(F, (tc, td, ta, tb, trho), grad_outputs=t64(gF))
        out = dict(c=c, t=t, d=d, a=a, b=b, rho=rho, gF=gF, F=F.detach().numpy(), gc=gc.numpy(),
                   gd=gd.numpy(), ga=ga.numpy(), gb=gb.numpy(), grho=grho.numpy())
        if single_too:
            # single-model twin on model 0 with the per-sample Jacobian the sensitivity notebook builds
            td, ta, tb, trho, tc = (t64(x[0], True) for x in (d, a, b, rho, c))
            F1 = ref_vec_gpu.dltar_vector(tc, t64(t[0]), td, ta, tb, trho, 2, -1, device="cpu")
            assert torch.equal(F1.detach(), F[0].detach())
            eye = torch.eye(N, dtype=torch.float64)
            J = [torch.autograd.grad(F1, x, grad_outputs=(eye,), retain_graph=True, is_grads_batched=True)[0].numpy()
                 for x in (td, ta, tb, trho, tc)]
            out.update(Jd=J[0], Ja=J[1], Jb=J[2], Jrho=J[3], Jc=J[4])
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, F.shape, float(F.abs().max()))


def case_grid_batched(name, model, S, vlist, tlist, seed, sigma=0.05):
    rng = np.random.default_rng(seed)
    d, a, b, rho = ensemble(model, S, sigma, rng)
    V = np.tile(vlist, (S, 1))
    T = np.tile(tlist, (S, 1))
    gdet = rng.standard_normal((S, len(vlist), len(tlist)))
    with fp64_reference():
        td, ta, tb, trho = (t64(x, True) for x in (d, a, b, rho))
        det = ref_matall.dltar_matrix(t64(V), t64(T), td, ta, tb, trho, 2, -1, device="cpu")
        gd, ga, gb, grho = torch.autograd.grad(det, (td, ta, tb, trho), grad_outputs=t64(gdet), retain_graph=True)
        gd1, ga1, gb1, grho1 = torch.autograd.grad(det, (td, ta, tb, trho), grad_outputs=torch.ones_like(det))
    np.savez_compressed(os.path.join(OUT, name), vlist=V, tlist=T, d=d, a=a, b=b, rho=rho, gdet=gdet,
                        det=det.detach().numpy(), gd=gd.numpy(), ga=ga.numpy(), gb=gb.numpy(), grho=grho.numpy(),
                        gd_ones=gd1.numpy(), ga_ones=ga1.numpy(), gb_ones=gb1.numpy(), grho_ones=grho1.numpy())
    print(name, det.shape, float(det.abs().max()))


def case_dmf(name, model, S, N, clist, trange, crange, seed, *, initial_method, inversion_method,
             compress, normalized, compress_method, damp_vertical, damp_horizontal, vp_vs_ratio=2.45):
    """multi_cal_grad.forward + backward(ones) exactly as multi_inversion drives it (_model.py:590-592)."""
    rng = np.random.default_rng(seed)
    d, a, b, rho = ensemble(model, S, 0.03, rng)
    if initial_method == "Constant":
        a = b * vp_vs_ratio
        rho = np.full_like(b, 2.0)
    t = np.tile(np.exp(rng.uniform(np.log(trange[0]), np.log(trange[1]), (1, N))), (S, 1))
    c = np.tile(rng.uniform(crange[0], crange[1], (1, N)), (S, 1))
    C = np.tile(clist, (S, 1))
    with fp64_reference():
        calc = ref_model.multi_cal_grad(t64(d), t64(a), t64(b), t64(rho), t64(C), damp_vertical, damp_horizontal,
                                        compress, normalized, compress_method, inversion_method, initial_method,
                                        vp_vs_ratio, [], device="cpu")
        loss = calc(t64(c), t64(t))
        loss.backward(torch.ones_like(loss))
        gb = calc.b.grad.numpy().copy()
        gd = calc.d.grad.numpy().copy() if inversion_method == "vs-and-thick" else np.zeros_like(gb)
    np.savez_compressed(os.path.join(OUT, name), c=c, t=t, C=C, d=d, a=a, b=b, rho=rho, loss=loss.detach().numpy(),
                        gb=gb, gd=gd,
                        meta=json.dumps(dict(initial_method=initial_method, inversion_method=inversion_method,
                                             compress=compress, normalized=normalized, compress_method=compress_method,
                                             damp_vertical=damp_vertical, damp_horizontal=damp_horizontal,
                                             vp_vs_ratio=vp_vs_ratio)))
    print(name, loss.detach().numpy())


def case_dmf_single(name):
    """iter_cal_grad.forward (single model, config 2 at reduced N,K) in float64."""
    rng = np.random.default_rng(7)
    d = np.asarray(MODEL6["thick"])
    b = np.asarray(MODEL6["vs"]) * (1 + rng.normal(0, 0.03, 6))
    a, rho = brocher(b)
    dis = np.loadtxt(os.path.join(REF, "examples/data/00_increase_layer/input/disper.txt"))[::20]
    clist = np.arange(2, 3.2, 0.01)
    with fp64_reference():
        calc = ref_model.iter_cal_grad(t64(d), t64(a), t64(b), t64(rho), t64(clist), 0.002, True, True, "exp",
                                       "vs-and-thick", "Brocher", 2.45, [], device="cpu")
        loss = calc(t64(dis[:, 1]), t64(dis[:, 0]))
        loss.backward()
        gb, gd = calc.b.grad.numpy().copy(), calc.d.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, name), c=dis[:, 1], t=dis[:, 0], C=clist, d=d, a=a, b=b, rho=rho,
                        loss=loss.detach().numpy(), gb=gb, gd=gd)
    print(name, float(loss))


def case_shipped_fixtures(name):
    """Rows of the reference's shipped known-answer files (roots + config-2 trajectory head)."""
    base = os.path.join(REF, "examples/data")
    d0 = np.loadtxt(os.path.join(base, "00_increase_layer/input/disper.txt"))
    d2 = np.loadtxt(os.path.join(base, "02_FRIUL7W/input/disper.txt"))
    d1 = np.loadtxt(os.path.join(base, "01_sandwich_layer/input/disper_disba.txt"))
    d1_obs = np.loadtxt(os.path.join(base, "01_sandwich_layer/input/disper.txt"))
    mj = json.load(open(os.path.join(base, "00_increase_layer/output/data/model.json")))
    loss = np.asarray(mj["inv_process"]["loss"], dtype=np.float64)
    out = dict(
        disper_00=d0, disper_02=d2, disper_01=d1, disper_01_obs=d1_obs,
        c2_init_thick=np.asarray(mj["init_model"]["thick"]), c2_init_vs=np.asarray(mj["init_model"]["vs"]),
        c2_init_vp=np.asarray(mj["init_model"]["vp"]), c2_init_rho=np.asarray(mj["init_model"]["rho"]),
        c2_loss_head=loss[:12], c2_iter_vs_head=np.asarray(mj["inv_process"]["iter_vs"][:12]),
        c2_iter_thick_head=np.asarray(mj["inv_process"]["iter_thick"][:12]),
        c2_loss_len=np.asarray(len(loss)), c2_loss_min=np.nanmin(loss), c2_loss_argmin=np.nanargmin(loss),
        friul=np.loadtxt(os.path.join(base, "02_FRIUL7W/input/FRIUL7W-structure.txt")),
    )
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, {k: np.shape(v) for k, v in out.items()})


def case_surf96_roots(name):
    """Scalar root search (`_surf96.surf96`) with float64 numpy inputs: the N1 root oracle."""
    out = {}
    # w5 / w5s: a water layer (Vs = 0) over four solid layers; in w5s the water is slower than every solid
    water = {"thick": [0.3, 0.4, 0.5, 0.6, 0.0], "vs": [0.0, 1.0, 1.5, 2.0, 2.6]}
    for tag, model, tr, n, modes, dc in (("m6", MODEL6, (0.1, 5), 60, [0], 0.001),
                                         ("m3", MODEL3, (0.01, 0.2), 60, [0, 1, 2], 0.001),
                                         ("m20", friul(20), (0.05, 5), 40, [0], 0.01),
                                         ("w5", water, (0.1, 4), 30, [0, 1], 0.005),
                                         ("w5s", water, (0.1, 4), 30, [0], 0.005)):
        if tag.startswith("w5"):
            vp, rho = brocher(np.maximum(model["vs"], 0.5))
            vp[0], rho[0] = (1.5 if tag == "w5" else 0.6), 1.03
        else:
            vp, rho = brocher(model["vs"])
        t = log_wavelength_periods(tr[0], tr[1], n)
        d, b = np.asarray(model["thick"], float), np.asarray(model["vs"], float)
        for mode in modes:
            c = ref_surf.surf96(t, d, vp, b, rho, mode, 0, 2, dc)
            out[f"{tag}_mode{mode}_c"] = np.asarray(c, dtype=np.float64)
            c = ref_surf.surf96(t, d, vp, b, rho, mode, 0, 1, dc)  # Love waves (ifunc = 1)
            out[f"{tag}_love_mode{mode}_c"] = np.asarray(c, dtype=np.float64)
        out[f"{tag}_t"], out[f"{tag}_d"], out[f"{tag}_a"], out[f"{tag}_b"], out[f"{tag}_rho"] = t, d, vp, b, rho
        out[f"{tag}_dc"] = np.asarray(dc)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, list(out))


def main():
    # NB: gradient fixtures keep trial velocities off the layer velocities: at c == beta_m (or alpha_m)
    # the reference's autograd returns NaN/ill-conditioned values (sqrt'(0)), SURVEY.md §7 "Branch semantics".
    torch.manual_seed(0)
    case_grid_single("grid_single_m6", MODEL6, np.arange(2, 3.2, 0.02), log_wavelength_periods(0.1, 5, 24))
    case_grid_single("grid_single_m3", MODEL3, np.arange(0.2, 0.4, 0.004), log_wavelength_periods(0.01, 0.2, 24))
    case_grid_single("grid_single_m20", friul(20), np.arange(0.5, 5.0, 0.09), log_wavelength_periods(0.05, 5, 24))
    case_points_batched("points_m6", MODEL6, 3, 48, (2.0, 3.2), (0.1, 5), 11)
    case_points_batched("points_m3", MODEL3, 4, 40, (0.2, 0.4), (0.01, 0.2), 12, sigma=0.01)
    case_points_batched("points_m20", friul(20), 3, 64, (0.5, 5.0), (0.05, 5), 13, sigma=2 / 30)
    case_points_batched("points_m30", friul(30), 2, 48, (0.5, 5.0), (0.05, 5), 14, sigma=0.02)
    case_grid_batched("grid_m20", friul(20), 3, 0.5 + 4.5 * np.arange(0, 1024, 16) / 1024,
                      log_wavelength_periods(0.05, 5, 16), 21, sigma=2 / 30)
    case_grid_batched("grid_m6", MODEL6, 2, np.arange(2.0037, 3.2, 0.03), log_wavelength_periods(0.1, 5, 12), 22)
    case_grid_batched("grid_m3", MODEL3, 3, np.arange(0.2012, 0.4, 0.005), log_wavelength_periods(0.01, 0.2, 12), 23, sigma=0.01)
    case_dmf("dmf_exp_m6", MODEL6, 3, 30, np.arange(2, 3.2, 0.02), (0.1, 5), (2.3, 3.0), 31, initial_method="Brocher",
             inversion_method="vs-and-thick", compress=True, normalized=True, compress_method="exp",
             damp_vertical=0.002, damp_horizontal=0.0)
    case_dmf("dmf_log_m6", MODEL6, 3, 30, np.arange(2, 3.2, 0.02), (0.1, 5), (2.3, 3.0), 32, initial_method="Brocher",
             inversion_method="vs", compress=True, normalized=True, compress_method="log",
             damp_vertical=0.0, damp_horizontal=0.001)
    case_dmf("dmf_raw_m20", friul(20), 3, 40, np.arange(0.5, 5.0, 0.05), (0.05, 5), (2.0, 3.6), 33,
             initial_method="Brocher", inversion_method="vs", compress=False, normalized=True, compress_method="exp",
             damp_vertical=0.0, damp_horizontal=0.0)
    case_dmf("dmf_const_m3", MODEL3, 4, 30, np.arange(0.2, 0.4, 0.002), (0.01, 0.2), (0.22, 0.38), 34,
             initial_method="Constant", inversion_method="vs-and-thick", compress=True, normalized=True,
             compress_method="exp", damp_vertical=0.0, damp_horizontal=0.0)
    case_dmf("dmf_expnonorm_m6", MODEL6, 2, 20, np.arange(2, 3.2, 0.02), (0.1, 5), (2.3, 3.0), 35,
             initial_method="Brocher", inversion_method="vs", compress=True, normalized=False, compress_method="exp",
             damp_vertical=0.001, damp_horizontal=0.0)
    case_dmf_single("dmf_single_m6")
    case_shipped_fixtures("shipped_fixtures")
    case_surf96_roots("surf96_roots")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "surf96_roots":  # only the root-search fixture
        case_surf96_roots("surf96_roots")
    else:
        main()
