"""Root search (`adsurf_b200._surf96.surf96`, host code) against the reference's scalar solver.

Golden: tests/golden/surf96_roots.npz = `_surf96.surf96` of the reference with float64 numpy inputs
(SURVEY.md §8c "Root-search oracle") for the three example models, modes 0-2; and the shipped
disper*.txt fixtures (float32-generated, matched to ~3e-6).  north_star tolerance: 1e-6 km/s.
"""
import numpy as np
import pytest

from adsurf_b200._surf96 import DispersionError, surf96
from tests.golden_models import brocher_np
from tests.helpers import load

ROOT_TOL = 1e-6  # km/s


@pytest.mark.parametrize("tag,modes", [("m6", [0]), ("m3", [0, 1, 2]), ("m20", [0])])
def test_roots_match_reference_solver(tag, modes):
    g = load("surf96_roots")
    t, d, a, b, rho, dc = (g[f"{tag}_{k}"] for k in ("t", "d", "a", "b", "rho", "dc"))
    for mode in modes:
        c = surf96(t, d, a, b, rho, mode, 0, 2, float(dc))
        ref = g[f"{tag}_mode{mode}_c"]
        assert np.array_equal(c > 0, ref > 0)  # same periods carry the mode (zeros where it does not exist)
        assert np.max(np.abs(c - ref)) <= ROOT_TOL


def test_shipped_fundamental_curve_config1():
    """examples/data/00_increase_layer/input/disper.txt (config 1; float32 run of the reference)"""
    s = load("shipped_fixtures")
    rows = s["disper_00"][::10]
    b = np.array([2.39, 2.54, 2.7, 2.82, 2.93, 3.04])
    a, rho = brocher_np(b)
    c = surf96(rows[:, 0], np.array([0.4, 0.5, 0.5, 0.5, 0.5, 0.5]), a, b, rho, 0, 0, 2, 0.001)
    assert np.max(np.abs(c - rows[:, 1])) < 5e-6


def test_group_velocity_and_errors():
    g = load("surf96_roots")
    t, d, a, b, rho = (g[f"m6_{k}"] for k in ("t", "d", "a", "b", "rho"))
    u = surf96(t[::6], d, a, b, rho, 0, 1, 2, 0.005)
    c = surf96(t[::6], d, a, b, rho, 0, 0, 2, 0.005)
    assert np.all(u > 0) and np.all(u <= c + 1e-9)  # normal dispersion: group <= phase velocity
    with pytest.raises(NotImplementedError):
        surf96(t, d, a, b, rho, 0, 0, 3, 0.005)  # fast delta matrix: not provided


def test_love_waves_match_reference_solver():
    """ifunc = 1 (dltar1): the host twin against the reference's float64 roots, every fixture model and mode"""
    g = load("surf96_roots")
    for tag, modes in (("m6", (0,)), ("m3", (0, 1, 2)), ("m20", (0,)), ("w5", (0, 1)), ("w5s", (0,))):
        t, d, a, b, rho = (g[f"{tag}_{k}"] for k in ("t", "d", "a", "b", "rho"))
        for mode in modes:
            ref = g[f"{tag}_love_mode{mode}_c"]
            c = surf96(t, d, a, b, rho, mode, 0, 1, float(g[f"{tag}_dc"]))
            assert np.array_equal(c > 0, ref > 0) and np.max(np.abs(c - ref)) <= 1e-12


def test_water_layer_matches_reference_solver():
    """layer 0 with Vs = 0 (the reference's llw = 0): Rayleigh roots of the host twin against the reference's, for a
    water layer faster (w5) and slower (w5s: the search starts below the water's Vp) than the solid layers"""
    g = load("surf96_roots")
    for tag, modes in (("w5", (0, 1)), ("w5s", (0,))):
        t, d, a, b, rho = (g[f"{tag}_{k}"] for k in ("t", "d", "a", "b", "rho"))
        assert b[0] == 0.0
        for mode in modes:
            ref = g[f"{tag}_mode{mode}_c"]
            c = surf96(t, d, a, b, rho, mode, 0, 2, float(g[f"{tag}_dc"]))
            assert np.array_equal(c > 0, ref > 0) and np.max(np.abs(c - ref)) <= 1e-12
