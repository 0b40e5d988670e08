"""ORACLE (test infrastructure, not product code): bivariate normal probabilities.

The reference calls ``scipy.stats.mvn.mvnun(lower, upper, mean, cov)``
(/root/reference/risk_assessment/collision_probability.py:247).  That is an
un-vendored third-party dependency (``scipy~=1.5.0``, Fortran ``MVNDST`` by
A. Genz; /root/reference/requirements.txt).  For two dimensions MVNDST does not
use its randomised lattice rule; it short-circuits to the deterministic
routine ``BVNMVN`` which combines four evaluations of ``BVU`` (upper-tail
bivariate normal, Drezner-Wesolowsky as modified by Genz, "Numerical
computation of rectangular bivariate and trivariate normal and t
probabilities", Statistics and Computing 14 (2004) 251-260).

This file restates that published algorithm in NumPy (vectorised) so the
oracle travels to machines without the reference.  It is pinned in
``tests/test_oracle_bvn.py`` against scipy's own C++ port of the same routine
(``scipy.special._ufuncs._bivariate_normal_cdf``), against closed forms
(rho = 0, Phi2(0,0,rho) = 1/4 + asin(rho)/(2 pi)) and against the golden
vectors produced by the reference's own ``get_collision_probability_fast``
under shims (tests/golden/).

Parity status: *unpinned by the reference* (it has no tests); pinned by the
substitutes listed above.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import ndtr  # Phi, the role of MVNPHI in mvndst.f

TWOPI = 2.0 * math.pi

# Gauss-Legendre half-rules used by BVU: 6, 12 and 20 points on [-1, 1]
# (positive abscissae; the rule is symmetric).
_GL = {
    3: (
        np.array([0.1713244923791705, 0.3607615730481384, 0.4679139345726904]),
        np.array([0.9324695142031522, 0.6612093864662647, 0.2386191860831970]),
    ),
    6: (
        np.array([0.04717533638651177, 0.1069393259953183, 0.1600783285433464,
                  0.2031674267230659, 0.2334925365383547, 0.2491470458134029]),
        np.array([0.9815606342467191, 0.9041172563704750, 0.7699026741943050,
                  0.5873179542866171, 0.3678314989981802, 0.1252334085114692]),
    ),
    10: (
        np.array([0.01761400713915212, 0.04060142980038694, 0.06267204833410906,
                  0.08327674157670475, 0.1019301198172404, 0.1181945319615184,
                  0.1316886384491766, 0.1420961093183821, 0.1491729864726037,
                  0.1527533871307259]),
        np.array([0.9931285991850949, 0.9639719272779138, 0.9122344282513259,
                  0.8391169718222188, 0.7463319064601508, 0.6360536807265150,
                  0.5108670019508271, 0.3737060887154196, 0.2277858511416451,
                  0.07652652113349733]),
    ),
}


def _bvu_lowcorr(h, k, r, lg):
    """|r| < 0.925 branch of BVU with an ``lg``-point half rule."""
    w, x = _GL[lg]
    hk = h * k
    hs = (h * h + k * k) / 2.0
    asr = np.arcsin(r)
    acc = np.zeros_like(h)
    for i in range(lg):
        for sgn in (-1.0, 1.0):  # X(I) is stored negative in mvndst.f
            sn = np.sin(asr * (sgn * -x[i] + 1.0) / 2.0)
            acc = acc + w[i] * np.exp((sn * hk - hs) / (1.0 - sn * sn))
    return acc * asr / (2.0 * TWOPI) + ndtr(-h) * ndtr(-k)


def _bvu_highcorr(h, k, r, lg):
    """|r| >= 0.925 branch of BVU."""
    w, x = _GL[lg]
    h = h.copy()
    k = np.where(r < 0, -k, k)
    hk = h * k
    bvn = np.zeros_like(h)
    lt1 = np.abs(r) < 1.0
    if np.any(lt1):
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            as_ = (1.0 - r) * (1.0 + r)
            a = np.sqrt(as_)
            bs = (h - k) ** 2
            c = (4.0 - hk) / 8.0
            d = (12.0 - hk) / 16.0
            b0 = a * np.exp(-(bs / as_ + hk) / 2.0) * (
                1.0 - c * (bs - as_) * (1.0 - d * bs / 5.0) / 3.0 + c * d * as_ * as_ / 5.0
            )
            b = np.sqrt(bs)
            b1 = np.exp(-hk / 2.0) * math.sqrt(TWOPI) * ndtr(-b / a) * b * (
                1.0 - c * bs * (1.0 - d * bs / 5.0) / 3.0
            )
            b0 = np.where(hk > -160.0, b0 - b1, b0)
            a2 = a / 2.0
            for i in range(lg):
                xs = (a2 * (-x[i] + 1.0)) ** 2
                rs = np.sqrt(1.0 - xs)
                b0 = b0 + a2 * w[i] * (
                    np.exp(-bs / (2.0 * xs) - hk / (1.0 + rs)) / rs
                    - np.exp(-(bs / xs + hk) / 2.0) * (1.0 + c * xs * (1.0 + d * xs))
                )
                xs = as_ * (x[i] + 1.0) ** 2 / 4.0
                rs = np.sqrt(1.0 - xs)
                b0 = b0 + a2 * w[i] * np.exp(-(bs / xs + hk) / 2.0) * (
                    np.exp(-hk * (1.0 - rs) / (2.0 * (1.0 + rs))) / rs
                    - (1.0 + c * xs * (1.0 + d * xs))
                )
            b0 = -b0 / TWOPI
        bvn = np.where(lt1, b0, bvn)
    pos = bvn + ndtr(-np.maximum(h, k))
    neg = -bvn + np.maximum(0.0, ndtr(-h) - ndtr(-k))
    return np.where(r > 0, pos, neg)


def bvu(sh, sk, r):
    """Upper-tail bivariate normal P(X > sh, Y > sk), unit variances, correlation r.

    Restates ``BVU`` of Genz' mvndst.f (un-vendored; see module docstring).
    Vectorised over broadcastable arrays.
    """
    sh, sk, r = np.broadcast_arrays(
        np.asarray(sh, dtype=np.float64),
        np.asarray(sk, dtype=np.float64),
        np.asarray(r, dtype=np.float64),
    )
    shape = sh.shape
    h = sh.ravel().copy()
    k = sk.ravel().copy()
    rr = r.ravel().copy()
    out = np.empty_like(h)
    ar = np.abs(rr)
    lg = np.where(ar < 0.3, 3, np.where(ar < 0.75, 6, 10))
    for n in (3, 6, 10):
        m = (lg == n) & (ar < 0.925)
        if np.any(m):
            out[m] = _bvu_lowcorr(h[m], k[m], rr[m], n)
    m = ar >= 0.925
    if np.any(m):
        out[m] = _bvu_highcorr(h[m], k[m], rr[m], 10)
    return out.reshape(shape)


def rect_prob_std(xl, xu, yl, yu, r):
    """P(xl < X < xu, yl < Y < yu) for a standard bivariate normal.

    Restates ``BVNMVN`` for INFIN = (2, 2) (both limits finite): the
    four-corner inclusion-exclusion of BVU, in mvndst.f's order, not clamped.
    """
    return bvu(xl, yl, r) - bvu(xu, yl, r) - bvu(xl, yu, r) + bvu(xu, yu, r)


def mvnun(lower, upper, mean, cov):
    """Drop-in for ``scipy.stats.mvn.mvnun`` restricted to two dimensions.

    mvnun standardises with ``stdev = sqrt(diag(cov))`` and takes the correlation
    from the strict lower triangle (``covar(i, j)``, ``j < i``), then calls MVNDST.
    Returns ``(value, inform)`` like the Fortran wrapper.
    """
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    mean = np.asarray(mean, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64)
    s0 = math.sqrt(cov[0, 0])
    s1 = math.sqrt(cov[1, 1])
    r = cov[1, 0] / s1 / s0
    xl = (lower[0] - mean[0]) / s0
    xu = (upper[0] - mean[0]) / s0
    yl = (lower[1] - mean[1]) / s1
    yu = (upper[1] - mean[1]) / s1
    return float(rect_prob_std(xl, xu, yl, yu, r)), 0


def mvnun_fast(lower, upper, mean, cov):
    """Same contract as :func:`mvnun`, using scipy's compiled port of BVU when it is importable
    (scalar calls are then as cheap as the reference's Fortran routine); used for CPU baselines."""
    try:
        from scipy.special._ufuncs import _bivariate_normal_cdf as _bvu_c
    except Exception:  # pragma: no cover - scipy layout changed
        return mvnun(lower, upper, mean, cov)
    s0 = math.sqrt(cov[0][0])
    s1 = math.sqrt(cov[1][1])
    r = cov[1][0] / s1 / s0
    xl = (lower[0] - mean[0]) / s0
    xu = (upper[0] - mean[0]) / s0
    yl = (lower[1] - mean[1]) / s1
    yu = (upper[1] - mean[1]) / s1
    return float(_bvu_c(xl, yl, r) - _bvu_c(xu, yl, r) - _bvu_c(xl, yu, r) + _bvu_c(xu, yu, r)), 0
