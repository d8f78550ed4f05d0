"""Differentiable operators over the CUDA kernels (torch.autograd.Function wrappers).

    det_points(vc, t, d, alpha, beta, rho)      -> F   [S,N]   (or [N] for a single model)
    det_grid(vlist, tlist, d, alpha, beta, rho) -> det [S,K,NF] (or [K,NF])
    det_grid_minmax(...)                         -> (max_k det, min_k det) [S,NF], no autograd
    det_grid_signbits(...)                       -> packed sign bitmap of the dense grid
    dmf_data_misfit(vc, t, Clist, d, alpha, beta, rho, compress=...) -> loss [S]
    sensitivity(vc, t, d, alpha, beta, rho)      -> per-sample Jacobians (dict)

These replace `dltar_vector` / `dltar_matrix` of the reference (_cps/*.py) and the autograd tape
they build (≈10^3 ATen ops per layer): forward saves only the inputs, backward launches the
hand-written adjoint kernel, which recomputes the layer terms (SURVEY.md §8b "Autograd contract").
`autograd.grad(..., is_grads_batched=True)` (sensitivity notebook 02_2 cell 12) is served by the
per-sample Jacobian kernel: backward detects a batched upstream gradient and contracts it with the
Jacobian rows instead of launching the reducing adjoint.

Everything here needs CUDA; there is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _capi

COMPRESS = {
    "none": _capi.COMPRESS_NONE,
    "exp": _capi.COMPRESS_EXP,
    "log": _capi.COMPRESS_LOG,
    "exp_nonorm": _capi.COMPRESS_EXP_NONORM,
    "norm": _capi.COMPRESS_NORM,
}


def _c64(x):
    if x.dtype != torch.float64:
        x = x.to(torch.float64)
    return x.contiguous()


def _is_batched(t):
    """True for the BatchedTensors autograd hands to backward under is_grads_batched=True / vmap
    (they have no storage, so they cannot be passed to a kernel)."""
    try:
        t.data_ptr()
        return False
    except RuntimeError:
        return True


class _DetPoints(torch.autograd.Function):
    @staticmethod
    def forward(c, t, d, a, b, rho):
        return _capi.points_fwd(d, a, b, rho, t, c)

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(*inputs)

    @staticmethod
    def backward(ctx, gF):
        c, t, d, a, b, rho = ctx.saved_tensors
        if _is_batched(gF):
            # autograd.grad(..., is_grads_batched=True): one Jacobian-kernel launch, then contract the
            # (batched) upstream gradient with torch ops, which carry batching rules
            _, Jd, Ja, Jb, Jr, Jc = _capi.points_jac(d, a, b, rho, t, c)
            gd, ga, gb, gr = ((gF.unsqueeze(-1) * J).sum(dim=-2) for J in (Jd, Ja, Jb, Jr))
            return gF * Jc, None, gd, ga, gb, gr
        _, gd, ga, gb, gr, gc = _capi.points_bwd(d, a, b, rho, t, c, _c64(gF))
        return gc, None, gd, ga, gb, gr


class _DetGrid(torch.autograd.Function):
    @staticmethod
    def forward(c, t, d, a, b, rho):
        return _capi.grid_fwd(d, a, b, rho, t, c)[0]

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(*inputs)

    @staticmethod
    def backward(ctx, gdet):
        c, t, d, a, b, rho = ctx.saved_tensors
        if _is_batched(gdet):
            raise NotImplementedError("adsurf_b200: batched upstream gradients are supported by det_points only")
        _, gd, ga, gb, gr = _capi.grid_bwd(d, a, b, rho, t, c, _c64(gdet))
        return None, None, gd, ga, gb, gr


class _DmfMisfit(torch.autograd.Function):
    """Fused DMF data term: loss and its gradient come out of one kernel sequence in forward;
    backward only scales the stored per-model gradients by the upstream gradient."""

    @staticmethod
    def forward(c, t, C, d, a, b, rho, compress, need_grad):
        out = _capi.dmf_step(d, a, b, rho, t, c, C, compress, with_grad=need_grad)
        loss = out[0]
        grads = out[3:]
        if need_grad:
            return (loss,) + tuple(grads)
        z = loss.new_zeros(0)
        return loss, z, z, z, z

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.need_grad = inputs[8]
        ctx.save_for_backward(*output[1:])
        ctx.mark_non_differentiable(*output[1:])

    @staticmethod
    def backward(ctx, gloss, *unused):
        if not ctx.need_grad:
            raise RuntimeError("adsurf_b200: dmf_data_misfit was evaluated without gradients")
        gd, ga, gb, gr = ctx.saved_tensors
        w = gloss.reshape(-1, 1)
        return None, None, None, w * gd, w * ga, w * gb, w * gr, None, None


def _prep(x, device):
    x = torch.as_tensor(x)
    if x.device != device or x.dtype != torch.float64:
        x = x.to(device=device, dtype=torch.float64)
    return x


def _device_of(*xs, device=None):
    if device is not None:
        dev = torch.device(str(device).lower()) if not isinstance(device, torch.device) else device
        if dev.type == "cuda":
            return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())
    for x in xs:
        if torch.is_tensor(x) and x.is_cuda:
            return x.device
    if not torch.cuda.is_available():
        raise _capi.AdsurfError("adsurf_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def det_points(vc, t, d, alpha, beta, rho, device=None):
    """Secular function at (period, velocity) picks. Reference: dltar_vector(..., ifunc=2, llw=-1)."""
    dev = _device_of(vc, t, d, alpha, beta, rho, device=device)
    vc, t, d, alpha, beta, rho = (_prep(x, dev) for x in (vc, t, d, alpha, beta, rho))
    single = d.dim() == 1
    if single:
        vc, t = vc.reshape(1, -1), t.reshape(1, -1)
        d, alpha, beta, rho = (x.reshape(1, -1) for x in (d, alpha, beta, rho))
    if vc.shape != t.shape:
        vc, t = torch.broadcast_tensors(vc, t)
    F = _DetPoints.apply(vc.contiguous(), t.contiguous(), d.contiguous(), alpha.contiguous(), beta.contiguous(),
                         rho.contiguous())
    return F[0] if single else F


def _grid_args(vlist, tlist, d, alpha, beta, rho, device):
    dev = _device_of(vlist, tlist, d, alpha, beta, rho, device=device)
    vlist, tlist, d, alpha, beta, rho = (_prep(x, dev) for x in (vlist, tlist, d, alpha, beta, rho))
    single = d.dim() == 1
    if single:
        vlist, tlist = vlist.reshape(1, -1), tlist.reshape(1, -1)
        d, alpha, beta, rho = (x.reshape(1, -1) for x in (d, alpha, beta, rho))
    S = d.shape[0]
    vlist = vlist.reshape(S, -1) if vlist.dim() != 2 else vlist
    tlist = tlist.reshape(S, -1) if tlist.dim() != 2 else tlist
    return single, tuple(x.contiguous() for x in (vlist, tlist, d, alpha, beta, rho))


def det_grid(vlist, tlist, d, alpha, beta, rho, device=None):
    """Secular function on the (trial velocity x period) grid. Reference: dltar_matrix(..., 2, -1)."""
    single, args = _grid_args(vlist, tlist, d, alpha, beta, rho, device)
    det = _DetGrid.apply(*args)
    return det[0] if single else det


@torch.no_grad()
def det_grid_minmax(vlist, tlist, d, alpha, beta, rho, device=None):
    """(max over trial velocities, min over trial velocities) of the dense grid, never materialised."""
    single, (c, t, d, a, b, rho) = _grid_args(vlist, tlist, d, alpha, beta, rho, device)
    _, cmax, cmin, _ = _capi.grid_fwd(d, a, b, rho, t, c, want_det=False, want_minmax=True)
    return (cmax[0], cmin[0]) if single else (cmax, cmin)


@torch.no_grad()
def det_grid_signbits(vlist, tlist, d, alpha, beta, rho, device=None):
    """Packed sign bitmap int32 [S,K,ceil(NF/32)]: bit (f%32) of word f//32 set where det < 0."""
    single, (c, t, d, a, b, rho) = _grid_args(vlist, tlist, d, alpha, beta, rho, device)
    bits = _capi.grid_fwd(d, a, b, rho, t, c, want_det=False, want_signbits=True)[3]
    return bits[0] if single else bits


def dmf_data_misfit(vc, t, Clist, d, alpha, beta, rho, compress="exp", device=None):
    """Data term of the Determinant Misfit Function, per model: mean_i |G(F_i / n_i)|.

    Reference: multi_cal_grad.forward up to F_all (_model.py:448-464).  Differentiable w.r.t.
    d, alpha, beta, rho (the normalisation n is a constant, as under the reference's no_grad).
    """
    dev = _device_of(vc, t, d, alpha, beta, rho, device=device)
    vc, t, d, alpha, beta, rho = (_prep(x, dev).contiguous() for x in (vc, t, d, alpha, beta, rho))
    mode = COMPRESS[compress] if isinstance(compress, str) else int(compress)
    C = _prep(Clist, dev).contiguous() if Clist is not None else None
    need = torch.is_grad_enabled() and any(x.requires_grad for x in (d, alpha, beta, rho))
    return _DmfMisfit.apply(vc, t, C, d, alpha, beta, rho, mode, need)[0]


@torch.no_grad()
def sensitivity(vc, t, d, alpha, beta, rho, device=None):
    """Per-sample Jacobians of the secular function and the phase-velocity sensitivity kernels.

    Returns dict with F [N], dF_d{thick,vp,vs,rho} [N,L], dF_dc [N] and the implicit-function
    kernels dc_d* = -(dF/dm)/(dF/dc) [N,L]  (examples/02_2 cell 12 of the reference).  Batched
    inputs ([S,N], [S,L]) return a leading S axis.
    """
    dev = _device_of(vc, t, d, alpha, beta, rho, device=device)
    vc, t, d, alpha, beta, rho = (_prep(x, dev) for x in (vc, t, d, alpha, beta, rho))
    single = d.dim() == 1
    if single:
        vc, t = vc.reshape(1, -1), t.reshape(1, -1)
        d, alpha, beta, rho = (x.reshape(1, -1) for x in (d, alpha, beta, rho))
    F, Jd, Ja, Jb, Jr, Jc = _capi.points_jac(d.contiguous(), alpha.contiguous(), beta.contiguous(), rho.contiguous(),
                                             t.contiguous(), vc.contiguous())
    out = {"F": F, "dF_dthick": Jd, "dF_dvp": Ja, "dF_dvs": Jb, "dF_drho": Jr, "dF_dc": Jc}
    for k in ("thick", "vp", "vs", "rho"):
        out["dc_d" + k] = -out["dF_d" + k] / Jc.unsqueeze(-1)
    if single:
        out = {k: v[0] for k, v in out.items()}
    return out
