"""Monte-Carlo ensemble generator + pre-screen — mirror of the reference's ADsurf/_MonteCarlo.py.

`ADsurf_MonteCarlo(...)` keeps the reference signature (_MonteCarlo.py:17-32) and its random-number
call order, so with the same `np.random.seed` it returns the same ensemble.  The pre-screen
(`forward=True`), which the reference runs as `sampling_num` sequential single-model forward calls
(_MonteCarlo.py:175-191), is ONE batched launch here: `adsurf_dmf_step(compress=NORM,
with_grad=0)` evaluates mean_i |F_i / (max_k det - min_k det)| for all models at once.

`generator="device"`: the ensemble itself is drawn on the GPU (`adsurf_mc_generate`: Philox normals,
clip, depth perturbation, Brocher / Constant / USArray) and returned — with the pre-screen misfits — as CUDA tensors, so a
4096-model ensemble never touches the host; `inversion(...)` accepts those tensors.  Same distribution as the numpy
path, different random stream (seeded by `seed`); the numpy path stays the default for seeded parity with the reference.
"""
from __future__ import annotations

import numpy as np

from . import ops
from ._utils import gen_model, gen_model1


def _perturb_depths(thick, d, depth_down, depth_up, sigma_thick, sampling_num, layering_method, sampling_method):
    """depth perturbation of _MonteCarlo.py:100-139 (LR / LN layering)"""
    depth_up, depth_down = np.asarray(depth_up, dtype=float), np.asarray(depth_down, dtype=float)
    if depth_up[0] == 0:
        depth_up = depth_up[1:]
    if depth_down[0] == 0:
        depth_down = depth_down[1:]
    depth = np.cumsum(d)
    mc_depth = thick  # the reference perturbs in place
    for i in range(d.shape[0]):
        if sampling_method == "normal":
            draw = np.random.normal(loc=depth[i], scale=(depth_up[i] - depth_down[i]) / sigma_thick, size=sampling_num)
            if layering_method == "LR":
                mc_depth[:, i] = np.clip(draw, a_min=depth_down[i], a_max=depth_up[i])
            else:  # "LN": each depth at least wmin/3 below the previous one (:126-131)
                mc_depth[:, i] = draw
                lo = depth_down[i] if i == 0 else mc_depth[:, i - 1] + depth_down[i]
                mc_depth[:, i] = np.clip(thick[:, i], a_min=lo, a_max=depth_up[i])
    mc_depth = np.insert(mc_depth, 0, 0, axis=1)
    out = np.diff(mc_depth, axis=1)
    out = np.clip(out, a_min=depth_down[1], a_max=None)  # thickness >= wmin/3
    out[0] = d
    out[:, -1] = out[:, -2]
    return out


def ADsurf_MonteCarlo(pvs_obs, clist, d, a, b, rho, layering_method="", depth_up_boundary=[], depth_down_boundary=[],
                      vsrange_sign="mul", vsrange=[-0.2, 0.2], gen_velMethod="Brocher", vp_vs_ratio=2.45,
                      sampling_num=200, sampling_method="normal", sigma_vs=10, sigma_thick=0, ifunc=2, AK135_data=[],
                      forward=True, device=None, generator="numpy", seed=0):
    d, a, b, rho = (np.asarray(x, dtype=np.float64) for x in (d, a, b, rho))
    if generator == "device":
        return _device_ensemble(pvs_obs, clist, d, a, b, rho, layering_method, depth_up_boundary, depth_down_boundary,
                                vsrange_sign, vsrange, gen_velMethod, vp_vs_ratio, sampling_num, sampling_method, sigma_vs,
                                sigma_thick, ifunc, AK135_data, forward, device, seed)
    if vsrange_sign == "mul":
        vs_lo, vs_hi = b * vsrange[0], b * vsrange[1]
    else:
        vs_lo, vs_hi = b + vsrange[0], b + vsrange[1]
    mc_vs = np.ones((sampling_num, b.shape[0])) * b
    mc_thick = np.ones_like(mc_vs) * d
    mc_vp = np.ones_like(mc_vs) * a
    mc_rho = np.ones_like(mc_vs) * rho
    for i in range(b.shape[0]):  # one draw of sampling_num-1 values per layer, as the reference
        if sampling_method == "normal":
            draw = np.random.normal(loc=b[i], scale=(vs_hi[i] - vs_lo[i]) / sigma_vs, size=sampling_num - 1)
            mc_vs[1:, i] = np.clip(draw, vs_lo[i], vs_hi[i])
    if (len(depth_up_boundary) == d.shape[0] and len(depth_down_boundary) == d.shape[0] and sigma_thick > 0
            and layering_method in ("LR", "LN")):
        mc_thick = _perturb_depths(mc_thick, d, depth_down_boundary, depth_up_boundary, sigma_thick, sampling_num,
                                   layering_method, sampling_method)
    if gen_velMethod == "Brocher":
        _, vp, _, rr = gen_model(mc_thick[1:], mc_vs[1:], area=True)
        mc_vp[1:], mc_rho[1:] = vp, rr
    elif gen_velMethod == "Constant":
        _, vp, _, rr = gen_model1(mc_thick[1:], mc_vs[1:], vp_vs_ratio=vp_vs_ratio, rho=rho, area=True)
        mc_vp[1:], mc_rho[1:] = vp, rr
    elif gen_velMethod == "USArray":
        ak = np.asarray(AK135_data)
        for i in range(1, sampling_num):
            _, vp, vs, rr = gen_model(mc_thick[i], mc_vs[i], area=True)
            for j in np.nonzero(vs > 4.6)[0]:
                vp_j, rho_j = vp[j], rr[j]
                vp[j] = ak[np.argmin(np.abs(ak[:, 2] - vp_j))][2]
                rr[j] = ak[np.argmin(np.abs(ak[:, 1] - rho_j))][1]
            mc_vp[i], mc_rho[i] = vp, rr
    loss = np.zeros(sampling_num)
    if forward:
        if ifunc != 2:
            raise NotImplementedError("adsurf_b200 implements ifunc=2 (Rayleigh, Dunkin matrix) only")
        import torch
        obs = np.asarray(pvs_obs, dtype=np.float64)
        t = np.tile(obs[:, 0].reshape(1, -1), (sampling_num, 1))
        c = np.tile(obs[:, 1].reshape(1, -1), (sampling_num, 1))
        C = np.tile(np.asarray(clist, dtype=np.float64).reshape(1, -1), (sampling_num, 1))
        with torch.no_grad():
            loss = ops.dmf_data_misfit(c, t, C, mc_thick, mc_vp, mc_vs, mc_rho, compress="norm",
                                       device=device).cpu().numpy()
    return {"model": {"thick": mc_thick, "vs": mc_vs, "vp": mc_vp, "rho": mc_rho}, "loss": loss}


def _device_ensemble(pvs_obs, clist, d, a, b, rho, layering_method, depth_up, depth_down, vsrange_sign, vsrange,
                     gen_velMethod, vp_vs_ratio, sampling_num, sampling_method, sigma_vs, sigma_thick, ifunc, AK135_data,
                     forward, device, seed):
    """generation + pre-screen on the GPU; everything returned as CUDA tensors"""
    import torch

    from . import _capi
    if sampling_method != "normal":
        raise ValueError("sampling_method must be 'normal' (the only one the reference implements)")
    if gen_velMethod not in _capi.VP_MODE or _capi.VP_MODE[gen_velMethod] == 0:
        raise NameError("gen_velMethod must be one of Brocher / Constant / USArray")
    dev = ops._device_of(device=device)
    G = lambda x: torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=dev)
    if vsrange_sign == "mul":
        vs_lo, vs_hi = b * vsrange[0], b * vsrange[1]
    else:
        vs_lo, vs_hi = b + vsrange[0], b + vsrange[1]
    dep_lo = dep_hi = None
    layering = 0
    if (len(depth_up) == d.shape[0] and len(depth_down) == d.shape[0] and sigma_thick > 0
            and layering_method in ("LR", "LN")):
        up, down = np.asarray(depth_up, dtype=float), np.asarray(depth_down, dtype=float)
        up = up[1:] if up[0] == 0 else up          # the reference drops a leading zero (:101-104)
        down = down[1:] if down[0] == 0 else down
        if len(up) < d.shape[0] or len(down) < d.shape[0]:
            raise ValueError("depth boundaries shorter than the model after dropping the leading zero")
        dep_lo, dep_hi, layering = G(down[:d.shape[0]]), G(up[:d.shape[0]]), _capi.LAYERING[layering_method]
    ak = G(np.asarray(AK135_data, dtype=float).reshape(-1, 3)) if gen_velMethod == "USArray" and len(AK135_data) else None
    mc_d, mc_a, mc_b, mc_r = _capi.mc_generate(G(d), G(a), G(b), G(rho), G(vs_lo), G(vs_hi), sampling_num,
                                               vel_method=_capi.VP_MODE[gen_velMethod], vp_vs_ratio=vp_vs_ratio,
                                               sigma_vs=sigma_vs, sigma_thick=sigma_thick, layering=layering,
                                               depth_lo=dep_lo, depth_hi=dep_hi, ak135=ak, seed=seed)
    loss = torch.zeros(sampling_num, dtype=torch.float64, device=dev)
    if forward:
        if ifunc != 2:
            raise NotImplementedError("adsurf_b200 implements ifunc=2 (Rayleigh, Dunkin matrix) only")
        obs = np.asarray(pvs_obs, dtype=np.float64)
        t = G(obs[:, 0]).reshape(1, -1).expand(sampling_num, -1).contiguous()
        c = G(obs[:, 1]).reshape(1, -1).expand(sampling_num, -1).contiguous()
        C = G(np.asarray(clist, dtype=np.float64)).reshape(1, -1).expand(sampling_num, -1).contiguous()
        with torch.no_grad():
            loss = ops.dmf_data_misfit(c, t, C, mc_d, mc_a, mc_b, mc_r, compress="norm", device=dev)
    return {"model": {"thick": mc_d, "vs": mc_b, "vp": mc_a, "rho": mc_r}, "loss": loss}
