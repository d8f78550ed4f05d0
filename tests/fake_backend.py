"""CPU stand-in for adsurf_b200._capi used ONLY by the host-logic tests (autograd plumbing, vmap
rule, inversion loops, sharding).  It serves the binding's functions from the oracle so those
tests run without a GPU.  It is installed by monkeypatching inside tests; the product never sees it.
"""
import torch

from oracle import dltar_oracle as O


def _leaf(*xs):
    return tuple(x.detach().clone().requires_grad_(True) for x in xs)


def points_fwd(d, a, b, rho, t, c):
    return O.det_points(c, t, d, a, b, rho).detach()


@torch.enable_grad()
def points_bwd(d, a, b, rho, t, c, gF, want_F=False):
    d, a, b, rho, c = _leaf(d, a, b, rho, c)
    F = O.det_points(c, t, d, a, b, rho)
    gc, gd, ga, gb, gr = torch.autograd.grad(F, (c, d, a, b, rho), grad_outputs=gF)
    return (F.detach() if want_F else None), gd, ga, gb, gr, gc


@torch.enable_grad()
def points_jac(d, a, b, rho, t, c):
    S, N = c.shape
    L = d.shape[1]
    d, a, b, rho, c = _leaf(d, a, b, rho, c)
    F = O.det_points(c, t, d, a, b, rho)
    J = [torch.zeros(S, N, L, dtype=torch.float64) for _ in range(4)]
    Jc = torch.zeros(S, N, dtype=torch.float64)
    for s in range(S):
        for i in range(N):
            g = torch.autograd.grad(F[s, i], (d, a, b, rho, c), retain_graph=True)
            for k in range(4):
                J[k][s, i] = g[k][s]
            Jc[s, i] = g[4][s, i]
    return F.detach(), J[0], J[1], J[2], J[3], Jc


def grid_fwd(d, a, b, rho, t, c, want_det=True, want_minmax=False, want_signbits=False):
    det = O.det_grid(c, t, d, a, b, rho).detach()
    cmax = det.max(1).values if want_minmax else None
    cmin = det.min(1).values if want_minmax else None
    return (det if want_det else None), cmax, cmin, None


@torch.enable_grad()
def grid_bwd(d, a, b, rho, t, c, gdet=None, want_det=False, out=None, ws=None):
    d, a, b, rho = _leaf(d, a, b, rho)
    det = O.det_grid(c, t, d, a, b, rho)
    g = torch.autograd.grad(det, (d, a, b, rho), grad_outputs=torch.ones_like(det) if gdet is None else gdet)
    return ((det.detach() if want_det else None),) + tuple(g)


_MODE = {0: dict(compress=False), 1: dict(compress=True, normalized=True, compress_method="exp"),
         2: dict(compress=True, normalized=True, compress_method="log"),
         3: dict(compress=True, normalized=False, compress_method="exp"),
         4: dict(compress=True, normalized=True, compress_method="none")}


@torch.enable_grad()
def dmf_step(d, a, b, rho, t, c, C, compress, with_grad=True, want_F=False, want_norm=False):
    d, a, b, rho = _leaf(d, a, b, rho)
    loss = O.dmf_loss(c, t, C, d, b, a=a, rho=rho, initial_method="given", **_MODE[int(compress)])
    if with_grad:
        grads = torch.autograd.grad(loss, (d, a, b, rho), grad_outputs=torch.ones_like(loss), allow_unused=True)
        grads = tuple(torch.zeros_like(x) if g is None else g for g, x in zip(grads, (d, a, b, rho)))
    else:
        grads = (None,) * 4
    return (loss.detach(), None, None) + grads


def install(monkeypatch):
    """route adsurf_b200 through the oracle on the CPU (tests of host logic only)"""
    import adsurf_b200._capi as capi
    import adsurf_b200._model as model
    import adsurf_b200.ops as ops
    for name in ("points_fwd", "points_bwd", "points_jac", "grid_fwd", "grid_bwd", "dmf_step"):
        monkeypatch.setattr(capi, name, globals()[name])
    cpu = torch.device("cpu")
    monkeypatch.setattr(ops, "_device_of", lambda *xs, device=None: cpu)
    monkeypatch.setattr(model, "_dev", lambda device: cpu)
