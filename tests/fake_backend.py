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


def _vp(b):
    return 0.9409 + 2.0947 * b - 0.8206 * b**2 + 0.2683 * b**3 - 0.0251 * b**4


def _dvp(b):
    return 2.0947 - 2 * 0.8206 * b + 3 * 0.2683 * b**2 - 4 * 0.0251 * b**3


def _rho(a):
    return 1.6612 * a - 0.4721 * a**2 + 0.0671 * a**3 - 0.0043 * a**4 + 0.000106 * a**5


def _drho(a):
    return 1.6612 - 2 * 0.4721 * a + 3 * 0.0671 * a**2 - 4 * 0.0043 * a**3 + 5 * 0.000106 * a**4


def _lap(x, axis, row0=0, n_all=None, before=None, after=None):
    """second difference tridiag(-1, 2, -1) with corner entries 1 along `axis`; `before` / `after` are the rows
    next to a shard of a longer axis (global length n_all, shard starting at row0)"""
    x = x.movedim(axis, 0)
    n = x.shape[0]
    n_all = n if n_all is None else n_all
    zero = torch.zeros_like(x[:1])
    xm = torch.cat((zero if before is None else before.reshape(x[:1].shape), x[:-1]), dim=0)
    xp = torch.cat((x[1:], zero if after is None else after.reshape(x[:1].shape)), dim=0)
    out = 2.0 * x - xm - xp
    if n_all == 1:
        out = x.clone()
    else:
        if row0 == 0:
            out[0] = x[0] - xp[0]
        if row0 + n == n_all:
            out[-1] = x[-1] - xm[-1]
    return out.movedim(0, axis)


class InversionLoop:
    """torch-CPU restatement of adsurf_b200._capi.InversionLoop / the C ABI adsurf_inversion_step (same interface,
    same order of operations as the kernels k_ens_prepare / k_ens_adam), data misfit and gradients from the oracle."""

    def __init__(self, state, consts, t, c, C, *, compress, vp_mode, vp_vs_ratio, invert_thick, layering, proj_flags,
                 damp_vertical, damp_horizontal, lr, gamma, step_size, T, beta1=0.9, beta2=0.999, eps=1e-8,
                 row0=0, S_all=None, halo_fn=None, **unused):
        self.st, self.k, self.t, self.c, self.C = state, consts, t, c, C
        b = state["b"]
        for key in ("mb", "vb", "md", "vd"):
            state.setdefault(key, torch.zeros_like(b))
        S, L = b.shape
        self.o = dict(compress=compress, vp_mode=vp_mode, ratio=vp_vs_ratio, thick=bool(invert_thick), layering=layering,
                      flags=proj_flags, dv=damp_vertical, dh=damp_horizontal, lr=lr, gamma=gamma, step=step_size,
                      b1=beta1, b2=beta2, eps=eps, row0=row0, S_all=S if S_all is None else S_all)
        self.halo_fn = halo_fn
        self.loss = torch.zeros(S, dtype=torch.float64)
        self.loss_hist = torch.full((T, S), 10.0, dtype=torch.float64)
        self.iter_b_hist = torch.zeros((T, S, L), dtype=torch.float64)
        self.iter_d_hist = torch.zeros((T, S, L), dtype=torch.float64) if invert_thick else None
        self.j = 0
        self.done = 0
        self.kernels_per_iteration, self.unroll = 0, 8

    @property
    def counter(self):
        return torch.tensor([self.j, 0], dtype=torch.int32)

    def prepare(self, n):
        pass

    def run(self, n):
        for _ in range(int(n)):
            self._iteration()
        self.done += int(n)

    def _adam(self, p, g, m, v, lr, step):
        o = self.o
        m.copy_(m + (1.0 - o["b1"]) * (g - m))
        v.copy_(o["b2"] * v + (1.0 - o["b2"]) * g * g)
        denom = torch.sqrt(v) / (1.0 - o["b2"] ** step) ** 0.5 + o["eps"]
        return p - (lr / (1.0 - o["b1"] ** step)) * (m / denom)

    @torch.no_grad()
    def _iteration(self):
        o, st, k, j = self.o, self.st, self.k, self.j
        b, d, alpha, rho = st["b"], st["d"], st["alpha"], st["rho"]
        S, L = b.shape
        out = dmf_step(d, alpha, b, rho, self.t, self.c, self.C, o["compress"], with_grad=True)
        loss, (gd, ga, gb, gr) = out[0].clone(), out[3:]
        g = gb.clone()
        if o["vp_mode"] in (1, 3):
            low = torch.ones_like(b, dtype=torch.bool) if o["vp_mode"] == 1 else b <= 4.6
            g = torch.where(low, gb + (ga + gr * _drho(alpha)) * _dvp(b), gb)
        elif o["vp_mode"] == 2:
            g = gb + ga * o["ratio"]
        sv = sh = None
        if o["dv"] > 0:
            mv = o["dv"] * _lap(b, 1) / L
            loss += mv.abs().sum(1)
            sv = torch.sign(mv)
        if o["dh"] > 0:
            halo = self.halo_fn(b) if self.halo_fn is not None else None
            r0, Sa = o["row0"], o["S_all"]
            before, after = (halo[1], halo[2]) if halo is not None else (None, None)
            mh = o["dh"] * _lap(b, 0, r0, Sa, before, after) / L
            loss += mh.abs().sum(1)
            sh = torch.sign(mh)
            sh_before = sh_after = None
            if halo is not None:
                if r0 > 0:
                    sh_before = torch.sign(o["dh"] * _lap(torch.stack((halo[1], b[0])), 0, r0 - 1, Sa, halo[0], None)[0] / L)
                if r0 + S < Sa:
                    sh_after = torch.sign(o["dh"] * _lap(torch.stack((b[-1], halo[2])), 0, r0 + S - 1, Sa, None, halo[3])[1] / L)
        # projection
        x = b.clone()
        if o["flags"] & 2:
            x = torch.where(torch.isnan(x), k["b0"], x)
        x = torch.where(torch.isnan(x), x, torch.minimum(torch.maximum(x, k["lower_b"]), k["upper_b"]))
        if (o["flags"] & 1) and L > 1:
            ok = ~(torch.isnan(x[:, -1]) | torch.isnan(x[:, -2]))
            x[:, -1] = torch.where(ok, torch.maximum(x[:, -1], x[:, -2]), x[:, -1])
        if j < self.iter_b_hist.shape[0]:
            self.iter_b_hist[j] = x
        th = d.clone()
        if o["thick"]:
            depth = torch.cumsum(d, dim=1)
            if o["layering"] == 1:
                depth = torch.minimum(torch.maximum(depth, k["mindepth"]), k["maxdepth"])
            elif o["layering"] == 2:
                cols = []
                for l in range(L):
                    lo = k["mindepth"][:, l] if l == 0 else cols[-1] + k["mindepth"][:, l]
                    cols.append(torch.minimum(torch.maximum(depth[:, l], lo), k["maxdepth"][:, l]))
                depth = torch.stack(cols, dim=1)
            th = torch.diff(torch.cat((torch.zeros_like(depth[:, :1]), depth), dim=1), dim=1)
            th = torch.maximum(th, k["mindepth"].reshape(-1)[0].expand_as(th))
            if j < self.iter_b_hist.shape[0]:
                self.iter_d_hist[j] = th
        if sv is not None:
            g = g + o["dv"] / L * _lap(sv, 1)
        if sh is not None:
            g = g + o["dh"] / L * _lap(sh, 0, o["row0"], o["S_all"], sh_before, sh_after)
        lr = o["lr"]
        for _ in range(j // o["step"] if o["step"] > 0 else 0):
            lr *= o["gamma"]
        xn = self._adam(x, g, st["mb"], st["vb"], lr, j + 1)
        b.copy_(xn)
        if o["vp_mode"] in (1, 3):
            low = torch.ones_like(b, dtype=torch.bool) if o["vp_mode"] == 1 else xn <= 4.6
            an = torch.where(low, _vp(xn), alpha)
            rho.copy_(torch.where(low, _rho(an), rho))
            alpha.copy_(an)
        elif o["vp_mode"] == 2:
            alpha.copy_(xn * o["ratio"])
        if o["thick"]:
            d.copy_(self._adam(th, gd, st["md"], st["vd"], lr, j + 1))
        self.loss.copy_(loss)
        if j < self.loss_hist.shape[0]:
            self.loss_hist[j] = loss
        self.j += 1


def install(monkeypatch):
    """route adsurf_b200 through the oracle on the CPU (tests of host logic only)"""
    import adsurf_b200._capi as capi
    import adsurf_b200._model as model
    import adsurf_b200.ops as ops
    for name in ("points_fwd", "points_bwd", "points_jac", "grid_fwd", "grid_bwd", "dmf_step"):
        monkeypatch.setattr(capi, name, globals()[name])
    monkeypatch.setattr(capi, "InversionLoop", InversionLoop)
    monkeypatch.setattr(capi, "fused_available", lambda device: True)
    cpu = torch.device("cpu")
    monkeypatch.setattr(ops, "_device_of", lambda *xs, device=None: cpu)
    monkeypatch.setattr(model, "_dev", lambda device: cpu)
