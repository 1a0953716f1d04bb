"""Millard-2012 muscle curves: quintic-Bezier construction and tabulation.

The reference's muscles are ``Millard2012EquilibriumMuscle`` objects whose four
normalised characteristic curves (active force-length, force-velocity, passive
fibre force-length, tendon force-length) are built by OpenSim's
``SmoothSegmentedFunctionFactory`` from quintic Bezier "corner" segments.
OpenSim is not available, so the published
construction is restated here (PARITY UNPINNED against OpenSim itself; the
known-answer properties f_L(1)=1, f_V(0)=1, f_V(-1)=0, f_V(1)=1.4,
f_PE(1.7)=1, f_T(1.049)=1 are tested in tests/test_curves.py).

For the GPU the curves are sampled once on the host into uniform cubic-Hermite
tables (value, slope); every muscle shares the same four tables because all
muscles in the reference models use the default curve parameters
(reference data/02905/02905_PRE/scale/model_predictive.osim muscle blocks).
"""
from __future__ import annotations

import numpy as np


def _scale_curviness(c: float) -> float:
    return 0.1 + 0.8 * c


def _corner(x0, y0, dydx0, x1, y1, dydx1, c):
    """Six control points of one quintic Bezier 'corner' segment."""
    if abs(dydx0 - dydx1) > 1e-12:
        xc = (y1 - y0 - x1 * dydx1 + x0 * dydx0) / (dydx0 - dydx1)
    else:
        xc = 0.5 * (x0 + x1)
    yc = (xc - x1) * dydx1 + y1
    x0m, y0m = x0 + c * (xc - x0), y0 + c * (yc - y0)
    x1m, y1m = x1 + c * (xc - x1), y1 + c * (yc - y1)
    return (np.array([x0, x0m, x0m, x1m, x1m, x1]),
            np.array([y0, y0m, y0m, y1m, y1m, y1]))


_B5 = np.array([1.0, 5.0, 10.0, 10.0, 5.0, 1.0])


def _bez(p, u):
    u = np.asarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    for k in range(6):
        out = out + _B5[k] * p[k] * u ** k * (1 - u) ** (5 - k)
    return out


def _dbez(p, u):
    u = np.asarray(u, dtype=np.float64)
    d = 5.0 * (p[1:] - p[:-1])
    b4 = np.array([1.0, 4.0, 6.0, 4.0, 1.0])
    out = np.zeros_like(u)
    for k in range(5):
        out = out + b4[k] * d[k] * u ** k * (1 - u) ** (4 - k)
    return out


class BezierCurve:
    """C2 piecewise quintic-Bezier function y(x) with linear extrapolation."""

    def __init__(self, segs):
        self.segs = segs
        self.x0 = segs[0][0][0]
        self.x1 = segs[-1][0][-1]
        self.y0 = segs[0][1][0]
        self.y1 = segs[-1][1][-1]
        self.dydx0 = self._end_slope(segs[0], 0.0)
        self.dydx1 = self._end_slope(segs[-1], 1.0)

    @staticmethod
    def _end_slope(seg, u):
        dx = _dbez(seg[0], np.array([u]))[0]
        dy = _dbez(seg[1], np.array([u]))[0]
        return dy / dx

    def eval(self, x):
        """Return (y, dy/dx) at x (array), solving x(u)=x by bisection+Newton."""
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        y = np.empty_like(x)
        d = np.empty_like(x)
        lo_mask = x <= self.x0
        hi_mask = x >= self.x1
        y[lo_mask] = self.y0 + self.dydx0 * (x[lo_mask] - self.x0)
        d[lo_mask] = self.dydx0
        y[hi_mask] = self.y1 + self.dydx1 * (x[hi_mask] - self.x1)
        d[hi_mask] = self.dydx1
        for px, py in self.segs:
            m = (~lo_mask) & (~hi_mask) & (x >= px[0]) & (x <= px[-1])
            if not m.any():
                continue
            xs = x[m]
            a = np.zeros_like(xs)
            b = np.ones_like(xs)
            u = (xs - px[0]) / (px[-1] - px[0])
            for _ in range(80):
                f = _bez(px, u) - xs
                a = np.where(f < 0, u, a)
                b = np.where(f >= 0, u, b)
                df = _dbez(px, u)
                un = u - f / np.where(df != 0, df, 1.0)
                bad = (un <= a) | (un >= b) | (df == 0)
                u = np.where(bad, 0.5 * (a + b), un)
            y[m] = _bez(py, u)
            d[m] = _dbez(py, u) / _dbez(px, u)
            lo_mask = lo_mask | m  # each x handled once
        return y, d


def active_force_length_curve(x0=0.4441, x1=0.73, x2=1.0, x3=1.8123,
                              ylow=0.0, dydx=0.8616, curviness=1.0):
    """Millard active force-length curve (ylow = 0 because fibre damping is
    on, see model_predictive.osim ``<minimum_value>0``)."""
    c = _scale_curviness(curviness)
    x_delta = 0.05 * x2
    xs = x2 - x_delta
    y0 = 0.0
    y1 = 1 - dydx * (xs - x1)
    dydx01 = 1.25 * (y1 - y0) / (x1 - x0)
    x01 = x0 + 0.5 * (x1 - x0)
    y01 = y0 + 0.5 * (y1 - y0)
    x1s = x1 + 0.5 * (xs - x1)
    y1s = y1 + 0.5 * (1 - y1)
    y2, y3 = 1.0, 0.0
    x23 = (x2 + x_delta) + 0.5 * (x3 - (x2 + x_delta))
    y23 = y2 + 0.5 * (y3 - y2)
    dydx23 = (y3 - y2) / ((x3 - x_delta) - (x2 + x_delta))
    segs = [_corner(x0, ylow, 0.0, x01, y01, dydx01, c),
            _corner(x01, y01, dydx01, x1s, y1s, dydx, c),
            _corner(x1s, y1s, dydx, x2, y2, 0.0, c),
            _corner(x2, y2, 0.0, x23, y23, dydx23, c),
            _corner(x23, y23, dydx23, x3, ylow, 0.0, c)]
    return BezierCurve(segs)


def force_velocity_curve(fmax_e=1.4, dydx_c=0.0, dydx_near_c=0.25,
                         dydx_iso=5.0, dydx_e=0.0, dydx_near_e=0.15,
                         conc_curviness=0.6, ecc_curviness=0.9):
    cc = _scale_curviness(conc_curviness)
    ce = _scale_curviness(ecc_curviness)
    xc, yc = -1.0, 0.0
    xnc = -0.9
    ync = yc + 0.5 * dydx_near_c * (xnc - xc) + 0.5 * dydx_c * (xnc - xc)
    xiso, yiso = 0.0, 1.0
    xe, ye = 1.0, fmax_e
    xne = 0.9
    yne = ye + 0.5 * dydx_near_e * (xne - xe) + 0.5 * dydx_e * (xne - xe)
    segs = [_corner(xc, yc, dydx_c, xnc, ync, dydx_near_c, cc),
            _corner(xnc, ync, dydx_near_c, xiso, yiso, dydx_iso, cc),
            _corner(xiso, yiso, dydx_iso, xne, yne, dydx_near_e, ce),
            _corner(xne, yne, dydx_near_e, xe, ye, dydx_e, ce)]
    return BezierCurve(segs)


def fiber_force_length_curve(e_zero=0.0, e_iso=0.7, k_low=0.2, k_iso=None,
                             curviness=0.75):
    if k_iso is None:
        k_iso = 2.0 / (e_iso - e_zero)
    c = _scale_curviness(curviness)
    xz, yz = 1 + e_zero, 0.0
    xiso, yiso = 1 + e_iso, 1.0
    dx = min(0.1 * (1.0 / k_iso), 0.1 * (xiso - xz))
    xlow = xz + dx
    xfoot = xz + 0.5 * (xlow - xz)
    ylow = 0.0 + k_low * (xlow - xfoot)
    segs = [_corner(xz, yz, 0.0, xlow, ylow, k_low, c),
            _corner(xlow, ylow, k_low, xiso, yiso, k_iso, c)]
    return BezierCurve(segs)


def tendon_force_length_curve(e_iso=0.049, k_iso=None, f_toe=2.0 / 3.0,
                              curviness=0.5):
    if k_iso is None:
        k_iso = 1.375 / e_iso
    c = _scale_curviness(curviness)
    x0, y0 = 1.0, 0.0
    xiso, yiso = 1.0 + e_iso, 1.0
    ytoe = f_toe
    xtoe = (ytoe - 1) / k_iso + xiso
    xfoot = 1.0 + (xtoe - 1.0) / 10.0
    yfoot = 0.0
    ytoe_mid = ytoe * 0.5
    xtoe_mid = (ytoe_mid - yiso) / k_iso + xiso
    dydx_toe_mid = (ytoe_mid - yfoot) / (xtoe_mid - xfoot)
    xtoe_ctrl = xfoot + 0.5 * (xtoe_mid - xfoot)
    ytoe_ctrl = yfoot + dydx_toe_mid * (xtoe_ctrl - xfoot)
    segs = [_corner(x0, y0, 0.0, xtoe_ctrl, ytoe_ctrl, dydx_toe_mid, c),
            _corner(xtoe_ctrl, ytoe_ctrl, dydx_toe_mid, xtoe, ytoe, k_iso, c)]
    return BezierCurve(segs)


# order of the four curves inside the flat tables
CURVE_FL, CURVE_FV, CURVE_FPE, CURVE_FT = 0, 1, 2, 3
CURVE_NAMES = ("active_force_length", "force_velocity",
               "fiber_force_length", "tendon_force_length")


def default_curves():
    return (active_force_length_curve(), force_velocity_curve(),
            fiber_force_length_curve(), tendon_force_length_curve())


def tabulate(curve: BezierCurve, n_intervals: int):
    """Uniform cubic-Hermite table over [curve.x0, curve.x1].

    Returns (x0, x1, y0_end_slope, y1_end_slope, table[(n+1), 2]) where table
    rows are (y_i, h * dy/dx_i), h = (x1-x0)/n.  Outside the domain both the
    exact curve and the table extrapolate linearly with the end slopes.
    """
    xs = np.linspace(curve.x0, curve.x1, n_intervals + 1)
    y, d = curve.eval(xs)
    # end knots: use exact end values/slopes (eval() treats them as outside)
    h = (curve.x1 - curve.x0) / n_intervals
    tab = np.stack([y, d * h], axis=1)
    return curve.x0, curve.x1, curve.dydx0, curve.dydx1, tab


def hermite_eval(x0, x1, tab, x):
    """Reference evaluation of a tabulated curve (numpy); mirrors what the
    CUDA device function does.  Returns (y, dy/dx)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    n = tab.shape[0] - 1
    h = (x1 - x0) / n
    t = (x - x0) / h
    i = np.clip(np.floor(t).astype(np.int64), 0, n - 1)
    s = t - i
    y0, m0 = tab[i, 0], tab[i, 1]
    y1, m1 = tab[i + 1, 0], tab[i + 1, 1]
    s2, s3 = s * s, s * s * s
    y = (2 * s3 - 3 * s2 + 1) * y0 + (s3 - 2 * s2 + s) * m0 + \
        (-2 * s3 + 3 * s2) * y1 + (s3 - s2) * m1
    dy = ((6 * s2 - 6 * s) * y0 + (3 * s2 - 4 * s + 1) * m0 +
          (-6 * s2 + 6 * s) * y1 + (3 * s2 - 2 * s) * m1) / h
    lo = x < x0
    hi = x > x1
    y = np.where(lo, tab[0, 0] + tab[0, 1] / h * (x - x0), y)
    dy = np.where(lo, tab[0, 1] / h, dy)
    y = np.where(hi, tab[n, 0] + tab[n, 1] / h * (x - x1), y)
    dy = np.where(hi, tab[n, 1] / h, dy)
    return y, dy
