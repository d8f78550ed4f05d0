"""Synthetic workloads of BASELINE.json (config 5) — inputs only, no arithmetic of the hot path.

The base model is the FRIUL7W crustal structure the reference ships as
examples/data/02_FRIUL7W/input/FRIUL7W-structure.txt (thickness and Vs columns; public Earth-model
data, embedded here because /root/reference is absent on the GPU box).  The ensemble is drawn the
way the reference's own generator does (`ADsurf_MonteCarlo`, _MonteCarlo.py:81-98,140-146, as
called by notebook 02_1 cell 9): for every layer i, Vs[1:, i] ~ Normal(vs_i, (2/sigma_vs)) clipped
to vs_i +- 1, row 0 = the unperturbed model, thickness unchanged, Vp/rho by Brocher — with
`np.random.seed(seed)` and the same draw order, so the arrays are bit-identical to the reference's.
"""
from __future__ import annotations

import numpy as np

from ._utils import gen_model

FRIUL7W_THICK = np.array([0.057, 0.043, 0.2, 0.7, 2.0, 0.1, 0.2, 0.2, 1.0, 0.5, 0.5, 0.2, 0.2, 0.1, 4.5, 0.1, 0.1, 0.1,
                          0.7, 2.5, 5.0, 5.0, 1.0, 2.0, 2.0, 7.5, 4.0, 3.0, 1.5, 9.0, 1.0])
FRIUL7W_VS = np.array([0.6, 1.8, 2.5, 3.05, 3.24, 3.14, 3.1, 3.06, 3.03, 3.02, 3.03, 3.06, 3.1, 3.14, 3.25, 3.4, 3.5,
                       3.6, 3.75, 3.77, 3.8, 3.82, 3.82, 3.85, 3.85, 3.75, 3.85, 4.25, 4.5, 4.65, 4.65])


def friul7w(nlayers=20):
    """(thick, vp, vs, rho) of the first `nlayers` rows, Vp/rho from Brocher (notebook 02_0 cell 5)."""
    return gen_model(FRIUL7W_THICK[:nlayers].copy(), FRIUL7W_VS[:nlayers].copy(), area=True)


def config5_ensemble(S=4096, nlayers=20, seed=0, sigma_vs=30, vsrange=(-1.0, 1.0)):
    """d, alpha, beta, rho [S, nlayers] float64 of the config-5 Monte-Carlo ensemble."""
    thick, vp, vs, rho = friul7w(nlayers)
    lo, hi = vs + vsrange[0], vs + vsrange[1]
    b = np.ones((S, nlayers)) * vs
    state = np.random.get_state()
    np.random.seed(seed)
    try:
        for i in range(nlayers):
            b[1:, i] = np.clip(np.random.normal(loc=vs[i], scale=(hi[i] - lo[i]) / sigma_vs, size=S - 1), lo[i], hi[i])
    finally:
        np.random.set_state(state)
    d = np.ones_like(b) * thick
    _, a, b, r = gen_model(d, b, area=True)
    a[0], r[0] = vp, rho
    return d, a, b, r


def config5_periods(NF=128):
    """log-wavelength period sampling of the reference (_ADsurf.py:217-218) over [1/20 s, 5 s]."""
    return np.exp(np.linspace(np.log(1 / 20), np.log(5), NF))


def config5_velocities(K=1024):
    """trial phase velocities c_k = 0.5 + 4.5 k / K, the [0.5, 5) km/s span of notebook 02_0 cell 3."""
    return 0.5 + 4.5 * np.arange(K) / K


def config5(S=4096, NF=128, K=1024, nlayers=20, seed=0):
    d, a, b, rho = config5_ensemble(S, nlayers, seed)
    t = np.tile(config5_periods(NF), (S, 1))
    c = np.tile(config5_velocities(K), (S, 1))
    return d, a, b, rho, t, c
