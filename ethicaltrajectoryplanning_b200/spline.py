"""Reference-path spline as per-segment coefficient tables.

The reference builds ``CubicSpline2D(x, y)`` from the un-vendored
``commonroad_helper_functions`` (planner/planning.py:375-377) and only ever calls
``csp.calc_position(s)`` on the hot path (planner/Frenet/utils/frenet_functions.py:278).
The boundary conditions of that third-party spline are unknown, so the device code takes
*coefficient tables* (knots ``s_k`` and ``a + b ds + c ds^2 + d ds^3`` per segment).  This
class builds such tables (natural end conditions over cumulative chord length) and accepts
foreign spline objects that expose PythonRobotics-style ``s`` / ``sx.{a,b,c,d}`` / ``sy`` tables.
"""
from __future__ import annotations

import numpy as np


def _natural_coeffs(s, f):
    """Natural cubic spline coefficients (a, b, c, d) per segment for samples f(s)."""
    n = len(s)
    h = np.diff(s)
    if np.any(h <= 0):
        raise ValueError("spline knots must be strictly increasing")
    A = np.zeros((n, n))
    rhs = np.zeros(n)
    A[0, 0] = 1.0
    A[-1, -1] = 1.0
    for i in range(1, n - 1):
        A[i, i - 1] = h[i - 1]
        A[i, i] = 2.0 * (h[i - 1] + h[i])
        A[i, i + 1] = h[i]
        rhs[i] = 3.0 * ((f[i + 1] - f[i]) / h[i] - (f[i] - f[i - 1]) / h[i - 1])
    c = np.linalg.solve(A, rhs)
    a = f[:-1]
    b = (f[1:] - f[:-1]) / h - h * (2.0 * c[:-1] + c[1:]) / 3.0
    d = (c[1:] - c[:-1]) / (3.0 * h)
    return np.stack([a, b, c[:-1], d], axis=1)


class CubicSpline2D:
    """2-D cubic spline over arclength with ``calc_position`` like the reference's ``csp``."""

    def __init__(self, x=None, y=None, *, knots=None, coef_x=None, coef_y=None):
        if knots is not None:
            self.s = np.ascontiguousarray(knots, dtype=np.float64)
            self.coef_x = np.ascontiguousarray(coef_x, dtype=np.float64)
            self.coef_y = np.ascontiguousarray(coef_y, dtype=np.float64)
        else:
            x = np.asarray(x, dtype=np.float64)
            y = np.asarray(y, dtype=np.float64)
            self.s = np.concatenate([[0.0], np.cumsum(np.hypot(np.diff(x), np.diff(y)))])
            self.coef_x = _natural_coeffs(self.s, x)
            self.coef_y = _natural_coeffs(self.s, y)
        if self.coef_x.shape != (len(self.s) - 1, 4) or self.coef_y.shape != self.coef_x.shape:
            raise ValueError("coefficient tables must be [n_segments][4]")

    @classmethod
    def from_foreign(cls, csp):
        """Tables from a PythonRobotics-style object (``s``, ``sx.a/b/c/d``, ``sy.a/b/c/d``)."""
        if isinstance(csp, cls):
            return csp
        if hasattr(csp, "coef_x"):
            return cls(knots=csp.s, coef_x=csp.coef_x, coef_y=csp.coef_y)
        if hasattr(csp, "cx") and hasattr(csp, "cy"):
            return cls(knots=csp.s, coef_x=csp.cx, coef_y=csp.cy)
        n = len(csp.s) - 1

        def tab(sp):
            return np.stack([np.asarray(sp.a)[:n], np.asarray(sp.b)[:n],
                             np.asarray(sp.c)[:n], np.asarray(sp.d)[:n]], axis=1)

        return cls(knots=np.asarray(csp.s), coef_x=tab(csp.sx), coef_y=tab(csp.sy))

    def segment(self, s):
        return np.clip(np.searchsorted(self.s, s, side="right") - 1, 0, len(self.s) - 2)

    def calc_position(self, s):
        """x(s), y(s); outside the knot range the first / last segment is extrapolated."""
        s = np.asarray(s, dtype=np.float64)
        i = self.segment(s)
        ds = s - self.s[i]
        cx, cy = self.coef_x[i], self.coef_y[i]
        x = cx[..., 0] + cx[..., 1] * ds + cx[..., 2] * ds * ds + cx[..., 3] * ds * ds * ds
        y = cy[..., 0] + cy[..., 1] * ds + cy[..., 2] * ds * ds + cy[..., 3] * ds * ds * ds
        return x, y

    def calc_yaw(self, s):
        s = np.asarray(s, dtype=np.float64)
        i = self.segment(s)
        ds = s - self.s[i]
        cx, cy = self.coef_x[i], self.coef_y[i]
        dx = cx[..., 1] + 2.0 * cx[..., 2] * ds + 3.0 * cx[..., 3] * ds * ds
        dy = cy[..., 1] + 2.0 * cy[..., 2] * ds + 3.0 * cy[..., 3] * ds * ds
        return np.arctan2(dy, dx)
