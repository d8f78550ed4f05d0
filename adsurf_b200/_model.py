"""DMF loss modules and inversion drivers — host-side mirror of the reference's ADsurf/_model.py.

Same class names, constructor arguments and result dictionaries as the reference
(`iter_cal_grad` _model.py:236, `multi_cal_grad` :349, `iter_inversion` :12, `multi_inversion` :480)
so the callers in `_ADsurf.inversion(...)` switch by import.  What differs is underneath:

  * the pointwise determinant, the dense-grid normalisation (max - min over the trial
    velocities), the compression and the per-model misfit reduction — and their gradient — are one
    fused kernel sequence (`ops.dmf_data_misfit` -> C ABI `adsurf_dmf_step`) instead of ~3.4e4 ATen
    calls per step; the Brocher Vp/rho chain rule and the Tikhonov terms stay in torch autograd on
    tiny [S,L] tensors;
  * arithmetic is float64 (the reference casts everything to float32, _utils.py:493-500);
  * the per-iteration clipping of Vs and of the layer depths runs on the device with vectorised
    torch ops (the reference round-trips through numpy and Python loops, _model.py:126-147,603-621).

Ordering inside one iteration is the reference's: loss -> backward -> clip parameters -> record
iterate -> optimizer.step() -> scheduler.step() (_model.py:118-162, :589-636).
"""
from __future__ import annotations

import numpy as np
import torch

import os

from . import _capi, _dist, ops
from ._utils import gen_model, gen_model1, ifunc_list, list2numpy, numpy2tensor, tensor2numpy

try:  # progress bars are optional
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(x, **kw):
        return x


def _dev(device):
    if isinstance(device, torch.device):
        dev = device
    else:
        dev = torch.device(str(device).lower())
    if dev.type != "cuda":
        # the reference's device="cpu" run: tensors live on the CPU, arithmetic still runs on the GPU
        return torch.device("cuda", torch.cuda.current_device())
    return dev


def brocher_vp(b):
    return 0.9409 + 2.0947 * b - 0.8206 * b**2 + 0.2683 * b**3 - 0.0251 * b**4


def brocher_rho(a):
    return 1.6612 * a - 0.4721 * a**2 + 0.0671 * a**3 - 0.0043 * a**4 + 0.000106 * a**5


def _second_difference(n, device):
    """Tikhonov operator of _model.py:292-294 / :407-409: tridiag(-1, 2, -1) with corner entries 1."""
    L0 = 2.0 * torch.eye(n, dtype=torch.float64, device=device)
    if n > 1:
        i = torch.arange(n - 1, device=device)
        L0[i, i + 1] = -1.0
        L0[i + 1, i] = -1.0
    L0[0, 0] = 1.0
    L0[n - 1, n - 1] = 1.0
    return L0


def _compress_mode(compress, normalized, compress_method):
    """Map the reference's three switches (_model.py:450-462) to the kernel's compression id."""
    if not compress:
        return "none"
    if not normalized:
        return "exp_nonorm"
    if compress_method == "log":
        return "log"
    if compress_method == "exp":
        return "exp"
    return "norm"  # normalised but no recognised compression: F/n is used as is


class _cal_grad_base(torch.nn.Module):
    def _velocities(self):
        """Vp and density from Vs inside the graph (_model.py:306-319 / :431-446)."""
        b = self.b
        if self.initial_method == "Brocher":
            a = brocher_vp(b)
            rho = brocher_rho(a)
        elif self.initial_method == "Constant":
            a = b * self.vp_vs_ratio
            rho = self.rho
        elif self.initial_method == "USArray":
            # the reference's AK135 substitution for Vs > 4.6 writes into temporaries (no effect); what survives
            # is its masked in-place update (_model.py:440-446): Brocher where Vs <= 4.6, and entries above
            # 4.6 km/s keep whatever Vp / rho they last had (self.a / self.rho persist between calls)
            low = b <= 4.6
            a = torch.where(low, brocher_vp(b), self.a)
            rho = torch.where(low, brocher_rho(a), self.rho)
            self.a, self.rho = a.detach(), rho.detach()
        else:
            a, rho = self.a, self.rho
        return a, rho


class multi_cal_grad(_cal_grad_base):
    """Batched DMF: loss[S] for S models at once (reference _model.py:349-478)."""

    def __init__(self, d, a, b, rho, Clist, damp_vertical, damp_horizontal, compress, normalized, compress_method,
                 inversion_method, initial_method, vp_vs_ratio, AK135_data=[], device="cpu"):
        super().__init__()
        self.device = _dev(device)
        t64 = lambda x: numpy2tensor(x).to(self.device)
        self.d = t64(d)
        self.a = t64(a)
        self.rho = t64(rho)
        self.b = torch.nn.Parameter(t64(b).clone())
        if inversion_method.lower() == "vs-and-thick":
            self.d = torch.nn.Parameter(self.d.clone())
        self.Clist = t64(Clist).contiguous()
        self.ifunc = ifunc_list["dunkin"]["rayleigh"]
        self.llw = 0 if float(self.b.detach()[0][0]) <= 0.0 else -1
        if self.llw == 0:
            raise NotImplementedError("adsurf_b200: water layer (Vs[0] <= 0) is not supported (broken in the reference too)")
        self.damp_vertical = damp_vertical
        self.damp_horizontal = damp_horizontal
        self.L_vertical = _second_difference(self.b.shape[1], self.device)
        self.L_horizontal = _second_difference(self.b.shape[0], self.device)
        self.shard = None      # (lo, hi, S) when this module holds rows lo:hi of a sharded ensemble
        self.coupling = None   # horizontal-damping rows owned by other ranks (added to the backward seed)
        self.normalized = normalized
        self.compress = compress
        self.compress_method = compress_method
        self.initial_method = initial_method
        self.vp_vs_ratio = vp_vs_ratio
        self.AK135_data = numpy2tensor(list2numpy(AK135_data)).to(self.device)

    def forward(self, vlist, tlist):
        a, rho = self._velocities()
        mode = _compress_mode(self.compress, self.normalized, self.compress_method)
        F_all = ops.dmf_data_misfit(vlist, tlist, self.Clist, self.d, a, self.b, rho, compress=mode,
                                    device=self.device)
        out = F_all
        nL = self.b.shape[1]
        if self.damp_vertical > 0:
            mv = self.damp_vertical * torch.matmul(self.b, self.L_vertical.T) / nL
            out = out + torch.sum(torch.abs(mv), dim=1)
        if self.damp_horizontal > 0:
            if self.shard is None:
                mh = self.damp_horizontal * torch.matmul(self.L_horizontal, self.b) / nL
                out = out + torch.sum(torch.abs(mh), dim=1)
            else:
                # sharded ensemble: the second difference across models needs the neighbours' Vs.
                # All ranks hold the same full operator; the other ranks' rows enter as constants and
                # the rows owned elsewhere are returned through `self.coupling` so that backward
                # still sees d(sum_r |L_h b|_r)/d(b_local) for every r this rank touches.
                lo, hi, S = self.shard
                with torch.no_grad():
                    b_all = _dist.all_gather_rows(self.b.detach(), S)
                b_full = torch.cat((b_all[:lo], self.b, b_all[hi:]), dim=0)
                mh = torch.sum(torch.abs(self.damp_horizontal * torch.matmul(self.L_horizontal, b_full) / nL), dim=1)
                out = out + mh[lo:hi]
                self.coupling = torch.sum(mh[:lo]) + torch.sum(mh[hi:])
        return out


class iter_cal_grad(_cal_grad_base):
    """Single-model DMF: scalar loss (reference _model.py:236-344)."""

    def __init__(self, d, a, b, rho, Clist, damp, compress, normalized, compress_method, inversion_method,
                 initial_method, vp_vs_ratio, AK135_data=[], device="cpu"):
        super().__init__()
        self.device = _dev(device)
        t64 = lambda x: numpy2tensor(x).to(self.device)
        self.d = t64(d).reshape(-1)
        self.a = t64(a).reshape(-1)
        self.rho = t64(rho).reshape(-1)
        self.b = torch.nn.Parameter(t64(b).reshape(-1).clone())
        if inversion_method.lower() == "vs-and-thick":
            self.d = torch.nn.Parameter(self.d.clone())
        self.Clist = t64(Clist).reshape(1, -1).contiguous()
        self.ifunc = ifunc_list["dunkin"]["rayleigh"]
        self.llw = 0 if float(self.b.detach()[0]) <= 0.0 else -1
        if self.llw == 0:
            raise NotImplementedError("adsurf_b200: water layer (Vs[0] <= 0) is not supported (broken in the reference too)")
        self.damp = damp
        self.L = _second_difference(self.b.shape[0], self.device)
        self.normalized = normalized
        self.compress = compress
        self.compress_method = compress_method
        self.initial_method = initial_method
        self.vp_vs_ratio = vp_vs_ratio
        self.AK135_data = numpy2tensor(list2numpy(AK135_data)).to(self.device)

    def forward(self, vlist, tlist):
        a, rho = self._velocities()
        mode = _compress_mode(self.compress, self.normalized, self.compress_method)
        vlist = numpy2tensor(vlist).to(self.device).reshape(1, -1)
        tlist = numpy2tensor(tlist).to(self.device).reshape(1, -1)
        F_all = ops.dmf_data_misfit(vlist, tlist, self.Clist, self.d.reshape(1, -1), a.reshape(1, -1),
                                    self.b.reshape(1, -1), rho.reshape(1, -1), compress=mode, device=self.device)[0]
        if self.damp > 0:
            m_norm = self.damp * torch.matmul(self.L, self.b) / self.b.shape[0]
            return F_all + torch.sum(torch.abs(m_norm))
        return F_all


def _clip_depths(thick, layer_mindepth, layer_maxdepth, layering_method):
    """Depth/thickness projection of _model.py:133-147 (single) and :603-620 (batched), on device.

    thick, layer_mindepth, layer_maxdepth: [..., L].  Returns the projected thickness."""
    depth = torch.cumsum(thick, dim=-1)
    if layering_method == "LR":
        depth = torch.minimum(torch.maximum(depth, layer_mindepth), layer_maxdepth)
    elif layering_method == "LN":
        cols = []
        for ind in range(depth.shape[-1]):
            lo = layer_mindepth[..., ind] if ind == 0 else cols[-1] + layer_mindepth[..., ind]
            # torch.clip / np.clip semantics: max wins when min > max
            cols.append(torch.minimum(torch.maximum(depth[..., ind], lo), layer_maxdepth[..., ind]))
        depth = torch.stack(cols, dim=-1)
    zero = torch.zeros_like(depth[..., :1])
    new_thick = torch.diff(torch.cat((zero, depth), dim=-1), dim=-1)
    floor = layer_mindepth[..., :1] if layer_mindepth.dim() == 1 else layer_mindepth[:1, :1]
    return torch.maximum(new_thick, floor.expand_as(new_thick))


class iter_inversion:
    """Single-model inversion loop (reference _model.py:12-234)."""

    def __init__(self, model_param, inv_param, init_model, pvs_obs, vsrange_sign="mul", vsrange=[0.1, 2],
                 AK135_data=[], device="cuda"):
        self.init_model = init_model
        self.model_param = model_param
        self.inv_param = inv_param
        self.pvs_obs = pvs_obs
        self.vsrange_sign = vsrange_sign
        self.vsrange = vsrange
        self.AK135_data = AK135_data
        self.device = _dev(device)
        self.inv_model = {"vs": [], "vp": [], "rho": [], "thick": []}
        self.inv_process = {"iter_vs": [], "iter_thick": [], "loss": []}
        self._run()

    def _run(self):
        im = self.init_model.init_model
        dev = self.device
        t64 = lambda x: numpy2tensor(list2numpy(x)).to(dev)
        b, a, rho, d = t64(im["vs"]), t64(im["vp"]), t64(im["rho"]), t64(im["thick"])
        layer_mindepth, layer_maxdepth = t64(im["layer_mindepth"]), t64(im["layer_maxdepth"])
        obs = list2numpy(self.pvs_obs)
        tlist, vlist = t64(obs[:, 0].reshape(-1)), t64(obs[:, 1].reshape(-1))
        ip, mp = self.inv_param, self.model_param
        Clist = t64(list2numpy(mp.clist).reshape(1, -1))
        inversion_method = ip.inversion_method
        calculator = iter_cal_grad(d, a, b, rho, Clist, ip.damp_vertical, ip.compress, ip.normalized,
                                   ip.compress_method, inversion_method, mp.initialize_method, mp.vp_vs_ratio,
                                   self.AK135_data, device=dev)
        if ip.optimizer != "Adam":
            raise NameError("The input optimizer {} can not find in Pytorch".format(ip.optimizer))
        optimizer = torch.optim.Adam(calculator.parameters(), lr=ip.lr)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=ip.step_size, gamma=ip.gamma)
        iteration = ip.iteration
        iter_vs = torch.zeros((iteration, b.shape[0]), dtype=torch.float64, device=dev)
        iter_thick = torch.zeros((iteration, d.shape[0]), dtype=torch.float64, device=dev)
        loss_list = torch.full((iteration,), 10.0, dtype=torch.float64, device=dev)
        if self.vsrange_sign == "mul":
            lower_b, upper_b = self.vsrange[0] * b, self.vsrange[1] * b
        else:
            lower_b, upper_b = b + self.vsrange[0], b + self.vsrange[1]
        patience, eps, trigger_times = 50, 1e-4, 0
        vs_and_thick = inversion_method.lower() == "vs-and-thick"
        last = iteration
        prev = None
        fused = os.environ.get("ADSURF_FUSED_OPT", "1") != "0" and _capi.fused_available(dev)
        if fused:
            # Device-resident loop (S = 1): a CUDA graph of whole iterations replayed in chunks; the reference's
            # early stopping (:168-186) looks at every loss, so it is replayed on the host over each finished chunk —
            # one synchronisation per `check_every` iterations instead of one per iteration.  Iterations that ran
            # past the stopping point inside a chunk only wrote history rows, which are reset below.
            method = mp.initialize_method if mp.initialize_method in _capi.VP_MODE else ""
            with torch.no_grad():
                a_now, r_now = calculator._velocities()
            c64 = lambda x: x.detach().to(torch.float64).reshape(1, -1).contiguous().clone()
            state = {"b": c64(calculator.b), "d": c64(calculator.d), "alpha": c64(a_now), "rho": c64(r_now)}
            consts = {"b0": c64(calculator.b), "alpha0": c64(a), "rho0": c64(rho), "lower_b": c64(lower_b),
                      "upper_b": c64(upper_b), "mindepth": c64(layer_mindepth), "maxdepth": c64(layer_maxdepth)}
            mode = ops.COMPRESS[_compress_mode(ip.compress, ip.normalized, ip.compress_method)]
            # single-model loops: half-space rule only on the vs-and-thick branch, no NaN repair (:122-159)
            loop = _capi.InversionLoop(
                state, consts, tlist.reshape(1, -1).contiguous(), vlist.reshape(1, -1).contiguous(), calculator.Clist,
                compress=mode, vp_mode=_capi.VP_MODE[method], vp_vs_ratio=mp.vp_vs_ratio, invert_thick=vs_and_thick,
                layering=_capi.LAYERING.get(mp.layering_method, 0),
                proj_flags=_capi.PROJ_LAST_LAYER if vs_and_thick else 0, damp_vertical=ip.damp_vertical,
                damp_horizontal=0.0, lr=ip.lr, gamma=ip.gamma, step_size=ip.step_size, T=iteration)
            check_every = int(os.environ.get("ADSURF_CHECK_EVERY", "64"))
            stop, j0 = None, 0
            while j0 < iteration and stop is None:
                n = min(check_every, iteration - j0)
                loop.run(n)
                chunk = loop.loss_hist[j0:j0 + n, 0].cpu().numpy()  # synchronises
                for j in range(j0, j0 + n):
                    cur = float(chunk[j - j0])
                    if j > 0:
                        if cur > 0.6:
                            trigger_times = 0
                        if (cur > prev) or (abs(cur - prev) / abs(prev) < eps):
                            trigger_times += 1
                            if trigger_times >= patience:
                                stop = j
                        else:
                            trigger_times = 0
                    if stop is None and cur != cur:  # NaN: drop this iterate and stop (:180-186)
                        last, stop = j, j
                    if stop is not None:
                        print("Early Stopping in Iteration:{}".format(j))
                        break
                    prev = cur
                j0 += n
            loss_list, iter_vs = loop.loss_hist[:, 0], loop.iter_b_hist[:, 0]
            if vs_and_thick:
                iter_thick = loop.iter_d_hist[:, 0]
            if stop is not None and stop + 1 < iteration:  # rows written past the stopping point
                loss_list[stop + 1:] = 10.0
                iter_vs[stop + 1:] = 0.0
                iter_thick[stop + 1:] = 0.0
        for j in tqdm(range(0 if not fused else iteration, iteration), disable=not getattr(ip, "progress", False)):
            loss = calculator(vlist, tlist)
            optimizer.zero_grad()
            loss.backward()
            with torch.no_grad():
                pb = calculator.b
                pb.data = torch.minimum(torch.maximum(pb.data, lower_b), upper_b)
                if vs_and_thick:
                    pb.data[-1] = torch.maximum(pb.data[-1], pb.data[-2])  # half-space >= layer above (:129)
                    iter_vs[j] = pb.data
                    pd = calculator.d
                    pd.data = _clip_depths(pd.data, layer_mindepth, layer_maxdepth, mp.layering_method)
                    iter_thick[j] = pd.data
                else:
                    iter_vs[j] = pb.data
            optimizer.step()
            scheduler.step()
            cur = float(loss.detach())  # the reference syncs every iteration for its progress bar (:164)
            loss_list[j] = cur
            if j > 0:
                if cur > 0.6:
                    trigger_times = 0
                if (cur > prev) or (abs(cur - prev) / abs(prev) < eps):
                    trigger_times += 1
                    if trigger_times >= patience:
                        print("Early Stopping in Iteration:{}".format(j))
                        break
                else:
                    trigger_times = 0
            if cur != cur:  # NaN: drop this iterate and stop (:180-186)
                last = j
                print("Early Stopping in Iteration:{}".format(j))
                break
            prev = cur
        loss_np = loss_list[:last].cpu().numpy()
        iter_vs = iter_vs[:last].cpu().numpy()
        iter_thick = iter_thick[:last].cpu().numpy()
        best = int(np.argmin(loss_np))
        inv_thick = iter_thick[best] if vs_and_thick else tensor2numpy(list2numpy(im["thick"]))
        inv_vs = iter_vs[best]
        inv_vp, inv_vs, inv_rho = _final_model(mp, inv_thick, inv_vs, self.AK135_data)
        out = _as_result(ip)
        self.inv_model = {"thick": out(np.asarray(inv_thick)), "vp": out(inv_vp), "vs": out(inv_vs), "rho": out(inv_rho)}
        self.inv_process = {"iter_vs": out(iter_vs), "iter_thick": out(iter_thick), "loss": out(loss_np)}


def _halo_rows(b, row0, S_all):
    """Vs rows row0-2, row0-1, row0+S, row0+S+1 of a sharded ensemble (zeros outside [0, S_all)): every rank
    contributes its first and last two rows to one small all-gather (reference coupling: _model.py:471-473)."""
    import torch.distributed as dist
    rank, world = _dist.rank_world()
    edge = torch.cat((b[:2], b[-2:]), dim=0).contiguous()
    got = [torch.empty_like(edge) for _ in range(world)]
    dist.all_gather(got, edge)
    out = torch.zeros_like(edge)
    if rank > 0:
        out[0:2] = got[rank - 1][2:4]
    if rank < world - 1:
        out[2:4] = got[rank + 1][0:2]
    return out


def _fused_ensemble_loop(self, calculator, vlist, tlist, Clist, a_init, rho_init, lower_b, upper_b, mindepth, maxdepth,
                         vs_and_thick, shard=None):
    """Iterations of multi_inversion with everything on the device: a CUDA graph of whole
    iterations (`adsurf_inversion_step`: dense pass, pointwise forward + adjoint with the DMF seed, velocity chain,
    Tikhonov terms, projection, Adam, StepLR, history) replayed `iteration` times with no host work in between.
    Same order of operations as the torch loop below (loss -> gradient -> projection -> record -> Adam -> StepLR);
    tests compare the two trajectories.  Sharded ensembles with horizontal damping exchange four Vs rows per
    iteration (`_halo_rows`) and issue the iterations one by one.  Returns (loss [T,S], iter_vs, iter_thick)."""
    ip, mp = self.inv_param, self.model_param
    c64 = lambda x: x.detach().to(torch.float64).contiguous().clone()
    b, d = c64(calculator.b), c64(calculator.d)
    method = mp.initialize_method if mp.initialize_method in _capi.VP_MODE else ""
    with torch.no_grad():  # Vp / rho of the first evaluation come from the velocity chain, like forward() does
        calculator_a, calculator_r = calculator._velocities()
    state = {"b": b, "d": d, "alpha": c64(calculator_a), "rho": c64(calculator_r)}
    full = lambda x: (torch.ones_like(b) * x).contiguous()
    consts = {"b0": c64(calculator.b), "alpha0": c64(a_init), "rho0": c64(rho_init), "lower_b": full(lower_b),
              "upper_b": full(upper_b), "mindepth": full(mindepth), "maxdepth": full(maxdepth)}
    mode = ops.COMPRESS[_compress_mode(ip.compress, ip.normalized, ip.compress_method)]
    row0, S_all, halo_fn = 0, b.shape[0], None
    if shard is not None:
        row0, _, S_all = shard
        if ip.damp_horizontal > 0:
            halo_fn = lambda bb: _halo_rows(bb, row0, S_all)
    kw = dict(compress=mode, vp_mode=_capi.VP_MODE[method], vp_vs_ratio=mp.vp_vs_ratio, invert_thick=vs_and_thick,
              layering=_capi.LAYERING.get(mp.layering_method, 0),
              proj_flags=_capi.PROJ_LAST_LAYER | (0 if vs_and_thick else _capi.PROJ_NAN_REPAIR),
              damp_vertical=ip.damp_vertical, damp_horizontal=ip.damp_horizontal, lr=ip.lr, gamma=ip.gamma,
              step_size=ip.step_size, T=ip.iteration)
    # models are independent without horizontal damping: large ensembles run as concurrent groups of rows, so that
    # one group's sweep fills the idle tail and the optimiser step of another (InversionLoopGroup)
    groups = _capi.loop_groups(b.shape[0], b.shape[1], vlist.shape[1], b.device, ip.damp_horizontal,
                                dense_pass=mode != ops.COMPRESS["none"])
    if groups > 1:
        loop = _capi.InversionLoopGroup(state, consts, tlist.contiguous(), vlist.contiguous(), Clist.contiguous(),
                                        groups=groups, **kw)
    else:
        loop = _capi.InversionLoop(state, consts, tlist.contiguous(), vlist.contiguous(), Clist.contiguous(),
                                   row0=row0, S_all=S_all, halo_fn=halo_fn, **kw)
    import time
    loop.prepare(ip.iteration)  # graph capture: set-up, not iterations
    if b.is_cuda:
        torch.cuda.synchronize(b.device)
    self._t_loop = time.perf_counter()
    loop.run(ip.iteration)
    with torch.no_grad():
        calculator.b.data.copy_(state["b"])
        if vs_and_thick:
            calculator.d.data.copy_(state["d"])
    self.loop = loop
    iter_thick = loop.iter_d_hist if vs_and_thick else loop.iter_b_hist[:0]  # Vs only: zeros, made by the caller
    return loop.loss_hist, loop.iter_b_hist, iter_thick


def _as_result(inv_param):
    """how result entries are returned: nested lists as in the reference (_model.py:691-701), or numpy arrays when
    inv_param.result_arrays is set"""
    if getattr(inv_param, "result_arrays", False):
        return lambda x: np.asarray(x, dtype=np.float64)
    return lambda x: np.asarray(x).tolist()


def _final_model(model_param, thick, vs, AK135_data):
    method = model_param.initialize_method
    if method == "Brocher":
        _, vp, vs, rho = gen_model(thick=thick, vs=vs, area=True)
    elif method == "Constant":
        _, vp, vs, rho = gen_model1(thick=thick, vs=vs, vp_vs_ratio=model_param.vp_vs_ratio, rho=model_param.rho,
                                    area=True)
    elif method == "USArray":
        _, vp, vs, rho = gen_model(thick=thick, vs=vs, area=True)
        ak = np.asarray(AK135_data)
        for j in np.nonzero(vs > 4.6)[0]:
            vp_j, rho_j = vp[j], rho[j]
            vp[j] = ak[np.argmin(np.abs(ak[:, 2] - vp_j))][2]
            rho[j] = ak[np.argmin(np.abs(ak[:, 1] - rho_j))][1]
    else:
        raise NameError("The initialize method can only select from [Brocher,Constant]")
    return vp, vs, rho


class multi_inversion:
    """Ensemble inversion loop: all models advance together (reference _model.py:480-701).

    With torch.distributed initialised (one process per GPU) the model axis is sharded across the
    ranks; the per-model losses and iterates are all-gathered at the end (and neighbouring rows
    of Vs every iteration when damp_horizontal > 0) — see adsurf_b200/_dist.py."""

    _fused_loop = _fused_ensemble_loop

    def __init__(self, model_param, inv_param, init_model, pvs_obs, vsrange_sign="mul", vsrange=[0.1, 2],
                 AK135_data=[], device="cuda"):
        self.init_model = init_model
        self.model_param = model_param
        self.inv_param = inv_param
        self.pvs_obs = pvs_obs
        self.vsrange_sign = vsrange_sign
        self.vsrange = vsrange
        self.AK135_data = AK135_data
        self.device = _dev(device)
        self.inv_model = {"vs": [], "vp": [], "rho": [], "thick": []}
        self.inv_process = {"iter_vs": [], "iter_thick": [], "loss": []}
        self._run()

    def _run(self):
        im = self.init_model.init_model
        dev = self.device
        t64 = lambda x: numpy2tensor(list2numpy(x)).to(dev)
        init_vs = list2numpy(im["vs"])
        b, a, rho, d = t64(init_vs), t64(im["vp"]), t64(im["rho"]), t64(im["thick"])
        S_all, L = b.shape
        obs = list2numpy(self.pvs_obs)
        tlist, vlist = t64(obs[:, :, 0]).contiguous(), t64(obs[:, :, 1]).contiguous()
        # one process per GPU: this rank owns models lo:hi of the ensemble
        sharded = _dist.is_sharded() and S_all >= _dist.rank_world()[1]
        lo, hi = _dist.shard_bounds(S_all) if sharded else (0, S_all)
        b_first = b[0].clone()
        b, a, rho, d, tlist, vlist = (x[lo:hi].contiguous() for x in (b, a, rho, d, tlist, vlist))
        S = hi - lo
        ip, mp = self.inv_param, self.model_param
        clist = t64(mp.clist).reshape(1, -1)
        Clist = clist.expand(S, -1).contiguous()
        inversion_method = ip.inversion_method
        layering_method = mp.layering_method
        layer_mindepth = torch.ones_like(b) * t64(im["layer_mindepth"])
        layer_maxdepth = torch.ones_like(b) * t64(im["layer_maxdepth"])
        calculator = multi_cal_grad(d, a, b, rho, Clist, ip.damp_vertical, ip.damp_horizontal, ip.compress,
                                    ip.normalized, ip.compress_method, inversion_method, mp.initialize_method,
                                    mp.vp_vs_ratio, self.AK135_data, device=dev)
        if sharded:
            calculator.shard = (lo, hi, S_all)
            calculator.L_horizontal = _second_difference(S_all, dev)
        if ip.optimizer != "Adam":
            raise NameError("The input optimizer {} can not find in Pytorch".format(ip.optimizer))
        optimizer = torch.optim.Adam(calculator.parameters(), lr=ip.lr)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=ip.step_size, gamma=ip.gamma)
        iteration = ip.iteration
        if self.vsrange_sign == "mul":
            lower_b, upper_b = self.vsrange[0] * b, self.vsrange[1] * b
        else:  # the reference offsets the FIRST model's Vs for every model (_model.py:577-578)
            lower_b = b_first * torch.ones_like(b) + self.vsrange[0]
            upper_b = b_first * torch.ones_like(b) + self.vsrange[1]
        vs_and_thick = inversion_method.lower() == "vs-and-thick"
        b0 = b.clone()
        ones = torch.ones(S, dtype=torch.float64, device=dev)
        # device-resident loop; a sharded ensemble with horizontal damping needs two rows per side from its neighbours
        fused = (os.environ.get("ADSURF_FUSED_OPT", "1") != "0" and _capi.fused_available(dev)
                 and not (sharded and ip.damp_horizontal > 0 and _dist.min_shard_rows(S_all) < 2))
        if not fused:
            iter_vs = torch.zeros((iteration, S, L), dtype=torch.float64, device=dev)
            iter_thick = torch.zeros((iteration, S, L), dtype=torch.float64, device=dev)
            loss_list = torch.full((iteration, S), 10.0, dtype=torch.float64, device=dev)
        import time
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t_setup = t_loop = time.perf_counter()
        if fused:
            # whole iterations on the device (adsurf_inversion_step in a CUDA graph), no torch autograd
            loss_list, iter_vs, iter_thick = self._fused_loop(
                calculator, vlist, tlist, Clist, a, rho, lower_b, upper_b, layer_mindepth, layer_maxdepth,
                vs_and_thick, shard=(lo, hi, S_all) if sharded else None)
            t_loop = self._t_loop
        for j in tqdm(range(0 if not fused else iteration, iteration), disable=not getattr(ip, "progress", False)):
            loss = calculator(vlist, tlist)
            optimizer.zero_grad()
            if calculator.coupling is not None:  # horizontal-damping rows owned by other ranks
                (loss.sum() + calculator.coupling).backward()
            else:
                loss.backward(ones)
            with torch.no_grad():
                pb = calculator.b
                if not vs_and_thick:  # NaN repair exists only on this branch in the reference (:624-626)
                    pb.data = torch.where(torch.isnan(pb.data), b0, pb.data)
                pb.data = torch.minimum(torch.maximum(pb.data, lower_b), upper_b)
                pb.data[:, -1] = torch.maximum(pb.data[:, -1], pb.data[:, -2])
                iter_vs[j] = pb.data
                if vs_and_thick:
                    pd = calculator.d
                    pd.data = _clip_depths(pd.data, layer_mindepth, layer_maxdepth, layering_method)
                    iter_thick[j] = pd.data
            optimizer.step()
            scheduler.step()
            loss_list[j] = loss.detach()
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        # seconds spent in the optimisation loop proper (what the reference's tqdm bar times); setup_s = building the
        # device-resident loop (buffers, CUDA-graph capture) before the first iteration
        t_end = time.perf_counter()
        self.timing = {"loop_s": t_end - t_loop, "setup_s": t_loop - t_setup, "iterations": iteration,
                       "fused": bool(fused), "gather_s": 0.0}
        loss_list = torch.where(torch.isnan(loss_list), torch.full_like(loss_list, 10.0), loss_list)
        if sharded:  # the exchange that ends the run: per-model losses and iterates of every rank
            loss_list = _dist.all_gather_rows(loss_list, S_all, dim=1)
            iter_vs = _dist.all_gather_rows(iter_vs, S_all, dim=1)
            if vs_and_thick:
                iter_thick = _dist.all_gather_rows(iter_thick, S_all, dim=1)
            S = S_all
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            self.timing["gather_s"] = time.perf_counter() - t_end
        loss_np = loss_list.cpu().numpy()
        iter_vs = iter_vs.cpu().numpy()
        # a Vs-only run records zeros as its thickness history (reference :609-621 never writes it): made on the host
        iter_thick = iter_thick.cpu().numpy() if vs_and_thick else np.zeros(iter_vs.shape)
        best_num = np.argmin(loss_np, axis=0)
        inv_thick = np.array(tensor2numpy(list2numpy(im["thick"])), dtype=np.float64, copy=True)
        if vs_and_thick:
            for i in range(S):
                inv_thick[i] = iter_thick[best_num[i], i, :]
        inv_vs, inv_vp, inv_rho = (np.zeros_like(inv_thick) for _ in range(3))
        for i in range(S):
            inv_vp[i], inv_vs[i], inv_rho[i] = _final_model(mp, inv_thick[i], iter_vs[best_num[i], i], self.AK135_data)
        out = _as_result(ip)
        self.inv_model = {"thick": out(inv_thick), "vp": out(inv_vp), "vs": out(inv_vs), "rho": out(inv_rho)}
        self.inv_process = {"iter_vs": out(iter_vs), "iter_thick": out(iter_thick), "loss": out(loss_np)}
        # host-side time after the loop (and the gather): device -> host copies of the histories, best-model pick,
        # result conversion
        self.timing["results_s"] = time.perf_counter() - t_end - self.timing["gather_s"]
