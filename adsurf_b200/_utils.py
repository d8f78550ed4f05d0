"""Host-side helpers with the names the reference's ADsurf/_utils.py exports.

Only cheap numpy/torch glue lives here (model parameterisation, Brocher relations, type
conversion); it keeps the semantics of the reference functions cited below so the inversion
drivers and the user API behave the same.  One deliberate difference: `numpy2tensor` produces
float64 (the reference casts everything to float32, _utils.py:493-500) because this build's
arithmetic is FP64.
"""
from __future__ import annotations

import warnings
from collections import namedtuple

import numpy as np
import torch

# reference _utils.py:20-23
ifunc_list = {
    "dunkin": {"love": 1, "rayleigh": 2},
    "fast-delta": {"love": 1, "rayleigh": 3},
}

DispersionCurve = namedtuple("DispersionCurve", ("period", "velocity", "mode", "wave", "type"))


class Dict(dict):
    """attribute-style dict (reference _utils.py:30-32)"""
    __setattr__ = dict.__setitem__
    __getattr__ = dict.__getitem__


def dictToObj(obj):
    if not isinstance(obj, dict):
        return obj
    out = Dict()
    for key, val in obj.items():
        out[key] = dictToObj(val)
    return out


# ------------------------------------------------------------------ conversions (_utils.py:493-531)
def numpy2tensor(a):
    if torch.is_tensor(a):
        return a.to(torch.float64)
    return torch.as_tensor(np.asarray(a, dtype=np.float64))


def tensor2numpy(a):
    if torch.is_tensor(a):
        return a.detach().cpu().numpy()
    return a


def list2numpy(a):
    return np.array(a) if isinstance(a, list) else a


def numpy2list(a):
    return a if isinstance(a, list) else a.tolist()


# ------------------------------------------------------------------ velocity models (_utils.py:45-249)
def brocher_vp_from_vs(vs):
    return 0.9409 + 2.0947 * vs - 0.8206 * vs**2 + 0.2683 * vs**3 - 0.0251 * vs**4


def brocher_rho_from_vp(vp):
    return 1.6612 * vp - 0.4721 * vp**2 + 0.0671 * vp**3 - 0.0043 * vp**4 + 0.000106 * vp**5


def _pack(thick, vp, vs, rho, area):
    if area:
        return thick, vp, vs, rho
    return {"thick": thick, "vp": vp, "vs": vs, "rho": rho}


def gen_model(thick, vs, area=False):
    """Brocher (2005) Vp and density from Vs (reference _utils.py:45-75)."""
    thick = np.array(thick)
    vs = np.array(vs)
    vp = brocher_vp_from_vs(vs)
    rho = brocher_rho_from_vp(vp)
    return _pack(thick, vp, vs, rho, area)


def gen_model1(thick, vs, vp_vs_ratio=2.45, rho=2, area=False):
    """Constant Vp/Vs ratio and density (reference _utils.py:76-110)."""
    thick = np.array(thick)
    vs = np.array(vs)
    return _pack(thick, vs * vp_vs_ratio, vs, np.ones_like(vs) * rho, area)


def _vs_from_dispersion(t, cg_obs, thick):
    """Initial Vs per layer from the observed curve (reference _utils.py:112-175): each layer takes
    the largest observed phase velocity whose 0.65-wavelength falls inside it, divided by 0.92; the
    half-space gets 1.1 x the largest observed velocity."""
    t = np.sort(np.asarray(tensor2numpy(t), dtype=np.float64).reshape(-1))
    cg = np.sort(np.asarray(tensor2numpy(cg_obs), dtype=np.float64).reshape(-1))
    thick = np.asarray(tensor2numpy(thick), dtype=np.float64).reshape(-1).copy()
    depth_of = 0.65 * t * cg
    n = len(thick)
    vs = np.zeros(n)
    top = 0.0
    for i in range(n - 1):
        if i > 0:
            top += thick[i - 1]
        bottom = top + thick[i]
        inside = (depth_of > top) & (depth_of < bottom)
        if inside.any():
            vs[i] = cg[inside].max() / 0.92
        else:
            vs[i] = cg[np.argmin(np.abs(depth_of - bottom))] / 0.92
    thick[n - 1] = 0
    vs[n - 1] = cg.max() * 1.1
    return thick, vs


def gen_init_model(t, cg_obs, thick, area=False):
    thick, vs = _vs_from_dispersion(t, cg_obs, thick)
    vp = brocher_vp_from_vs(vs)
    return _pack(thick, vp, vs, brocher_rho_from_vp(vp), area)


def gen_init_model1(t, cg_obs, thick, vp_vs_ratio=2.45, rho=2, area=False):
    thick, vs = _vs_from_dispersion(t, cg_obs, thick)
    return _pack(thick, vs * vp_vs_ratio, vs, np.ones_like(vs) * rho, area)


def model_rms_misfit(vs_true, vs_compare):
    vs_true, vs_compare = tensor2numpy(vs_true), tensor2numpy(vs_compare)
    return np.sum(vs_true - vs_compare) / len(vs_true)


def load_model(path):
    tab = np.loadtxt(path)
    return {"thick": tab[:, 0], "vp": tab[:, 1], "vs": tab[:, 2], "rho": tab[:, 3]}


# ------------------------------------------------------------------ layering (_utils.py:290-487)
def check_wavelengths(wmin, wmax):
    wmin, wmax = float(wmin), float(wmax)
    if wmin <= 0 or wmax <= 0:
        raise ValueError("Wavelength must be > 0.")
    if wmin > wmax:
        warnings.warn("Minimum wavelength must be less than maximum wavelength. Swapping!")
        wmin, wmax = wmax, wmin
    return wmin, wmax


def check_depth_factor(depth_factor):
    if not isinstance(depth_factor, (int, float)):
        raise TypeError(f"`factor` must be `int` or `float`. Not {type(depth_factor)}.")
    if depth_factor < 2:
        warnings.warn("`factor` must be >=2. Setting `factor` equal to 2.")
        depth_factor = 2
    return depth_factor


def check_layering_ratio(lr):
    if not isinstance(lr, (int, float)):
        raise TypeError(f"`lr` must be `int` or `float`, not {type(lr)}.")
    if lr <= 1:
        raise ValueError("`lr` must be greater than 1.")
    return lr


def depth_to_thick(depths):
    depths = np.array(depths)
    if depths[0] != 0:
        depths = np.insert(depths, 0, 0)
    return np.diff(depths)


def thick_to_depth(thicknesses):
    return [0] + [sum(thicknesses[:i]) for i in range(1, len(thicknesses))]


def depth_lr(wmin, wmax, lr, depth_factor):
    """Layering-Ratio depth bounds (Cox & Teague 2016; reference _utils.py:387-451)."""
    wmin, wmax = check_wavelengths(wmin, wmax)
    depth_factor = check_depth_factor(depth_factor)
    lr = check_layering_ratio(lr)
    lo, hi = [wmin / 3], [wmin / 2]
    dmax = wmax / depth_factor
    while hi[-1] < dmax:
        lo.append(hi[-1])
        step = hi[-1] if len(hi) == 1 else hi[-1] - hi[-2]
        hi.append(step * lr + hi[-1])
    if (dmax - hi[-2]) > (hi[-2] - hi[-3]):
        hi[-1] = dmax          # cap the deepest layer and open a half-space below it
        lo.append(dmax)
        hi.append(dmax + 1)
    else:
        hi[-2] = dmax          # stretch the previous layer, last entry becomes the half-space
        lo[-1] = dmax
        hi[-1] = dmax + 1
    return lo, hi


def depth_ln(wmin, wmax, nlayers, depth_factor=2):
    """Layering-by-Number depth bounds (reference _utils.py:453-487)."""
    wmin, wmax = check_wavelengths(wmin, wmax)
    if not isinstance(nlayers, int):
        raise TypeError(f"`nlayers` must be `int`. Not {type(nlayers)}.")
    if nlayers < 1:
        raise ValueError("Number of layers for must be >= 1.")
    depth_factor = check_depth_factor(depth_factor)
    return np.array([wmin / 3] * nlayers), np.array([wmax / depth_factor] * nlayers)
