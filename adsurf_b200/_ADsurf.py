"""User-facing API — mirror of the reference's ADsurf/_ADsurf.py (same names and call signatures).

    Model_param, inv_param                       configuration holders            (_ADsurf.py:22, :125)
    cal_determinant, cal_dispersion_curve        forward helpers                  (:201, :232)
    Model, True_model, Init_model                synthetic-test / initial models  (:315, :323, :381)
    init_model_MonteCarlo                        ensemble of initial models       (:541)
    inversion(...)                               single-model or ensemble DMF inversion (:651)

Host glue only: every determinant evaluation goes through the CUDA kernels (adsurf_b200.ops).
Plotting is not part of this package (the reference imports matplotlib/seaborn unconditionally at
_ADsurf.py:10; nothing on the compute path needs them).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._MonteCarlo import ADsurf_MonteCarlo
from ._model import iter_inversion, multi_inversion
from ._utils import (depth_ln, depth_lr, depth_to_thick, gen_init_model, gen_init_model1, ifunc_list, list2numpy,
                     numpy2list, numpy2tensor, tensor2numpy)

__all__ = ["Model_param", "inv_param", "cal_determinant", "cal_dispersion_curve", "Model", "True_model", "Init_model",
           "init_model_MonteCarlo", "inversion", "ADsurf_MonteCarlo"]


class Model_param:
    """Model / sampling parameters (reference _ADsurf.py:22-120)."""

    def __init__(self, dc=None, vmin=None, vmax=None, Nt=None, tmin=None, tmax=None, layering_method="",
                 initialize_method="", layering_ratio=None, depth_factor=None, layer_number=None, vp_vs_ratio=None,
                 rho=None, fundamental_range=[]):
        self.dc, self.vmin, self.vmax = dc, vmin, vmax
        self.Nt = Nt
        self.tmin, self.tmax = tmin, tmax
        self.clist = np.arange(self.vmin, self.vmax, self.dc)
        self.layering_method = layering_method
        self.layering_ratio = layering_ratio
        self.depth_factor = depth_factor
        self.layer_number = layer_number
        self.initialize_method = initialize_method
        self.vp_vs_ratio = vp_vs_ratio
        self.rho = rho
        self.fundamental_range = fundamental_range

    def _checkInput(self):
        if self.vmin > self.vmax:
            raise ValueError("the minimum phase velocity must less than the maximum phase velocity")
        if self.tmin > self.tmax:
            raise ValueError("the minimum period must less than the maximum period")
        if self.layering_method not in ["LN", "LR", "", None]:
            raise ValueError("you can take one of layering method including[LR/LN]")
        if len(self.fundamental_range) not in (0, 2):
            raise ValueError("The fundamental period range may set 2 number")


class inv_param:
    """Inversion parameters (reference _ADsurf.py:125-196).

    result_arrays (new keyword, default False = the reference's nested lists): return `inv_model` / `inv_process`
    entries as float64 numpy arrays — for a 1024-model, 400-iteration run the list conversion costs several times
    the optimisation loop itself."""

    def __init__(self, inversion_method="vs", wave="rayleigh", algorithm="dunkin", mode=0, compress=True,
                 compress_method="exp", normalized=True, lr=0, damp_vertical=0, damp_horizontal=0, iteration=200,
                 step_size=100, gamma=0.85, optimizer="Adam", result_arrays=False):
        self.result_arrays = result_arrays
        self.inversion_method = inversion_method
        self.wave = wave
        self.algorithm = algorithm
        self.mode = mode
        self.compress = compress
        self.compress_method = compress_method
        self.normalized = normalized
        self.lr = lr
        self.damp_vertical = damp_vertical
        self.damp_horizontal = damp_horizontal
        self.iteration = iteration
        self.step_size = step_size
        self.gamma = gamma
        self.optimizer = optimizer


def _periods(model_param, sampling_method, sampling_num):
    """period sampling of _ADsurf.py:209-218"""
    tmin, tmax = model_param.tmin, model_param.tmax
    if sampling_method == "frequency":
        return 1 / np.linspace(1 / tmax, 1 / tmin, sampling_num)[::-1]
    if sampling_method == "period":
        return np.linspace(tmin, tmax, sampling_num)
    if sampling_method == "log-wavelength":
        return np.exp(np.linspace(np.log(tmin), np.log(tmax), sampling_num))
    raise ValueError(f"unknown sampling_method {sampling_method!r}")


def cal_determinant(vel_model, model_param, sampling_method="log-wavelength", sampling_num=100, device=None):
    """Secular-function image F[K,NF] of one model on (arange(vmin,vmax,dc) x periods) (_ADsurf.py:201-230).

    Returns (tlist, vlist, F) as the reference does; F is a float64 tensor (on the CPU unless a CUDA
    `device` is requested), computed by the dense-grid kernel."""
    a, b, rho, d = (numpy2tensor(list2numpy(vel_model[k])) for k in ("vp", "vs", "rho", "thick"))
    tlist = numpy2tensor(_periods(model_param, sampling_method, sampling_num).copy())
    vlist = numpy2tensor(np.arange(model_param.vmin, model_param.vmax, model_param.dc))
    if ifunc_list["dunkin"]["rayleigh"] != 2 or float(b[0]) <= 0.0:
        raise NotImplementedError("adsurf_b200: Rayleigh / Dunkin without a water layer only")
    F = ops.det_grid(vlist, tlist, d, a, b, rho)
    want = torch.device("cpu") if device is None else torch.device(str(device).lower())
    return tlist, vlist, F.to(want)


def cal_dispersion_curve(vel_model, model_param, sampling_method="log-wavelength", sampling_num=100, dispOrder=[0],
                         device=None):
    """Modal dispersion curves [period, c, mode] by root search (_ADsurf.py:232-310 -> _surf96.surf96).

    device=None / "cpu": the sequential host solver (`_surf96.surf96`, as in the reference);
    device="cuda" (any capitalisation, or a torch.device): the same search run by the CUDA kernel
    (`dispersion.dispersion_curves`) — all requested modes in one launch, same mode bookkeeping, same roots."""
    from ._surf96 import surf96
    b, a, rho, d = (np.asarray(list2numpy(vel_model[k]), dtype=np.float64) for k in ("vs", "vp", "rho", "thick"))
    t = _periods(model_param, sampling_method, sampling_num)
    ifunc = ifunc_list["dunkin"]["rayleigh"]
    parts = []
    on_gpu = device is not None and torch.device(str(device).lower()).type == "cuda"
    if on_gpu:
        from .dispersion import dispersion_curves
        wanted = sorted(set(dispOrder))
        curves = dispersion_curves(t, d, a, b, rho, modes=wanted, dc=model_param.dc, device=device)
    for mode in sorted(set(dispOrder)):
        c = curves[wanted.index(mode)] if on_gpu else np.asarray(surf96(t, d, a, b, rho, mode, 0, ifunc, model_param.dc))
        ok = c > 0
        if mode > 0 and not ok.any():
            break  # the reference stops stacking at the first absent overtone
        parts.append(np.column_stack((t[ok], c[ok], np.full(int(ok.sum()), float(mode)))))
    return np.vstack(parts) if parts else np.zeros((0, 3))


class Model:
    def __init__(self):
        self.true_model = {}
        self.init_model = {}


class True_model:
    """A known model for synthetic tests (reference _ADsurf.py:323-375)."""

    def __init__(self, model_param, thick=[], vs=[], vp=[], rho=[]):
        self.thick, self.vp, self.vs, self.rho = (list2numpy(x) for x in (thick, vp, vs, rho))
        self.model_param = model_param
        self.tlist, self.vlist = [], []
        self.vel_model = {"vs": self.vs, "vp": self.vp, "rho": self.rho, "thick": self.thick}

    def _cal_determinant(self, sampling_method="log-wavelength", sampling_num=100):
        self.tlist, self.vlist, F = cal_determinant(self.vel_model, self.model_param, sampling_method, sampling_num)
        return F

    def _cal_dispersion_curve(self, sampling_method="log-wavelength", sampling_num=100, dispOrder=[0], device=None):
        return cal_dispersion_curve(self.vel_model, self.model_param, sampling_method, sampling_num, dispOrder, device)


class Init_model:
    """Parameterisation + initial model from the observed curve (reference _ADsurf.py:381-535)."""

    def __init__(self, model_param, pvs_obs=[], thick=[], vs=[], vp=[], rho=[]):
        self.pvs_obs = pvs_obs
        self.thick, self.vp, self.vs, self.rho = (list2numpy(x) for x in (thick, vp, vs, rho))
        self.model_param = model_param
        self.tlist, self.vlist = [], []
        self.init_model = {"vs": [], "vp": [], "rho": [], "thick": [], "layer_mindepth": [], "layer_maxdepth": []}
        self._checkDispData()
        self._parameterization()
        self._genInitModel()

    def _fundamental_mask(self):
        obs = self.pvs_obs
        mask = obs[:, 2] == 0 if obs.shape[1] == 3 else obs[:, 0] > 0
        fr = self.model_param.fundamental_range
        if len(fr) > 0:
            mask = mask * (obs[:, 0] > fr[0]) * (obs[:, 0] < fr[1])
        return mask

    def _parameterization(self):
        mp = self.model_param
        method = (mp.layering_method or "none").lower()
        obs = self.pvs_obs
        mask = self._fundamental_mask()
        wavelength = obs[:, 0][mask] * obs[:, 1][mask]
        wmin, wmax = np.min(wavelength), np.max(wavelength)
        if method == "lr":
            lo, hi = depth_lr(wmin=wmin, wmax=wmax, lr=mp.layering_ratio, depth_factor=mp.depth_factor)
            thick = depth_to_thick((np.array(lo) + np.array(hi)) / 2)
            thick[-1] = thick[-2]
            self.init_model["thick"] = np.array(thick)
        elif method == "ln":
            lo, hi = depth_ln(wmin=wmin, wmax=wmax, nlayers=mp.layer_number, depth_factor=mp.depth_factor)
            self.init_model["thick"] = np.array([hi[0] / mp.layer_number] * mp.layer_number)
        elif method in ("none", ""):
            if len(numpy2list(self.thick)) == 0:
                raise ValueError("No specify the layering method and not input the layer's parameter !!")
            self.init_model["thick"] = self.thick
            lo, hi = depth_ln(wmin=wmin, wmax=wmax, nlayers=len(self.thick), depth_factor=mp.depth_factor)
        else:
            raise ValueError(f"unknown layering_method {mp.layering_method!r}")
        self.init_model["layer_mindepth"] = lo
        self.init_model["layer_maxdepth"] = hi

    def _genInitModel(self):
        mp = self.model_param
        thick = list2numpy(self.init_model["thick"])
        obs = self.pvs_obs
        mask = obs[:, 2] == 0 if obs.shape[1] == 3 else obs[:, 0] > 0
        if mp.initialize_method == "Brocher":
            _, vp, vs, rho = gen_init_model(obs[:, 0][mask], obs[:, 1][mask], thick=thick, area=True)
        elif mp.initialize_method == "Constant":
            _, vp, vs, rho = gen_init_model1(obs[:, 0][mask], obs[:, 1][mask], thick=thick, vp_vs_ratio=mp.vp_vs_ratio,
                                             rho=mp.rho, area=True)
        else:
            vp, vs, rho = self.vp, self.vs, self.rho
            if len(numpy2list(vp)) == 0:
                raise NameError("No specify the initialize method and input vs/vp/rho values!!")
        self.init_model["vp"], self.init_model["vs"], self.init_model["rho"] = vp, vs, rho

    def _checkDispData(self):
        self.pvs_obs = np.asarray(list2numpy(self.pvs_obs), dtype=np.float64)
        if self.pvs_obs.ndim != 2:
            raise ValueError("The observed dispersion curve need a 2-D array")

    def _cal_determinant(self, sampling_method="log-wavelength", sampling_num=100):
        self.tlist, self.vlist, F = cal_determinant(self.init_model, self.model_param, sampling_method, sampling_num)
        return F

    def _cal_dispersion_curve(self, sampling_method="log-wavelength", sampling_num=100, dispOrder=[0], device=None):
        return cal_dispersion_curve(self.init_model, self.model_param, sampling_method, sampling_num, dispOrder, device)


class init_model_MonteCarlo:
    """Ensemble of initial models around `init_model` (reference _ADsurf.py:541-646)."""

    def __init__(self, model_param, init_model, sampling_num=200, sampling_method="normal", vsrange_sign="mul",
                 vsrange=[-0.2, 0.2], sigma_vs=10, sigma_thick=0, AK135_data=[], forward=True, generator="numpy",
                 seed=0, device=None):
        # generator="device": the ensemble is drawn and pre-screened on the GPU and stays there (CUDA tensors)
        self.generator, self.seed, self.device = generator, seed, device
        self.model_param = model_param
        self.init_model = init_model
        self.sampling_num = sampling_num
        self.sampling_method = sampling_method
        self.vsrange_sign = vsrange_sign
        self.vsrange = vsrange
        self.sigma_vs = sigma_vs
        self.sigma_thick = sigma_thick
        self.AK135_data = AK135_data
        self.forward = forward
        self.MonteCarlo_model = {"vs": [], "vp": [], "rho": [], "thick": [], "loss": []}
        self._MonteCarlo()

    def _MonteCarlo(self):
        im, mp = self.init_model.init_model, self.model_param
        res = ADsurf_MonteCarlo(
            self.init_model.pvs_obs, mp.clist, list2numpy(im["thick"]), list2numpy(im["vp"]), list2numpy(im["vs"]),
            list2numpy(im["rho"]), layering_method=mp.layering_method,
            depth_down_boundary=list2numpy(im["layer_mindepth"]), depth_up_boundary=list2numpy(im["layer_maxdepth"]),
            vsrange_sign=self.vsrange_sign, vsrange=self.vsrange, sampling_num=self.sampling_num,
            sampling_method=self.sampling_method, gen_velMethod=mp.initialize_method, vp_vs_ratio=mp.vp_vs_ratio,
            sigma_vs=self.sigma_vs, sigma_thick=self.sigma_thick, AK135_data=self.AK135_data, forward=self.forward,
            generator=self.generator, seed=self.seed, device=self.device)
        for key in ("vs", "vp", "rho", "thick"):
            self.MonteCarlo_model[key] = res["model"][key]
        self.MonteCarlo_model["loss"] = res["loss"]


def inversion(model_param, inv_param, init_model, pvs_obs, vsrange_sign="mul", vsrange=[0.1, 2], AK135_data=[],
              device="cuda"):
    """DMF inversion: one model (1-D initial Vs) or an ensemble (2-D) (reference _ADsurf.py:651-675).

    `device` accepts what the reference accepts plus any capitalisation ("Cuda", "cuda:0", "cpu",
    torch.device); the arithmetic always runs on the GPU."""
    vs0 = init_model.init_model["vs"]
    ndim = vs0.dim() if torch.is_tensor(vs0) else np.asarray(list2numpy(vs0)).ndim
    if ndim == 1:
        return iter_inversion(model_param, inv_param, init_model, pvs_obs, vsrange_sign, vsrange, AK135_data, device)
    if ndim == 2:
        return multi_inversion(model_param, inv_param, init_model, pvs_obs, vsrange_sign, vsrange, AK135_data, device)
    raise ValueError("init_model['vs'] must be 1-D (single model) or 2-D (ensemble)")
