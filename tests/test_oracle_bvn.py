"""Pins oracle/bvn.py (restated Genz BVU / MVNDST N=2) against scipy's port and closed forms."""
import numpy as np
import pytest
from scipy.special import ndtr

from oracle import bvn


def test_against_scipy_port_of_bvu():
    cdf = pytest.importorskip("scipy.special._ufuncs")._bivariate_normal_cdf  # upper-tail convention
    rng = np.random.default_rng(0)
    n = 100000
    h, k = rng.uniform(-7, 7, n), rng.uniform(-7, 7, n)
    r = rng.uniform(-0.999, 0.999, n)
    r[:3000] = rng.choice([0.0, 0.3, -0.3, 0.2999, 0.75, 0.7499, 0.925, -0.925, 0.9249, 0.9999, -0.9999], 3000)
    assert np.max(np.abs(bvn.bvu(h, k, r) - cdf(h, k, r))) < 5e-16


def test_closed_forms():
    r = np.linspace(-0.99, 0.99, 41)
    np.testing.assert_allclose(bvn.bvu(0.0, 0.0, r), 0.25 + np.arcsin(r) / (2 * np.pi), atol=2e-16)
    h, k = np.meshgrid(np.linspace(-4, 4, 9), np.linspace(-4, 4, 9))
    np.testing.assert_allclose(bvn.bvu(h, k, 0.0), ndtr(-h) * ndtr(-k), rtol=1e-14, atol=1e-300)


def test_rectangle_properties():
    # rho = 0: product of Phi differences; rectangle covering the plane: 1
    p = bvn.rect_prob_std(-0.5, 1.0, -2.0, 0.25, 0.0)
    np.testing.assert_allclose(p, (ndtr(1.0) - ndtr(-0.5)) * (ndtr(0.25) - ndtr(-2.0)), rtol=1e-14)
    np.testing.assert_allclose(bvn.rect_prob_std(-40.0, 40.0, -40.0, 40.0, 0.6), 1.0, rtol=1e-14)
    # additivity over a split rectangle
    a = bvn.rect_prob_std(-1.0, 0.3, -0.7, 1.1, -0.4)
    b = bvn.rect_prob_std(0.3, 2.0, -0.7, 1.1, -0.4)
    c = bvn.rect_prob_std(-1.0, 2.0, -0.7, 1.1, -0.4)
    np.testing.assert_allclose(a + b, c, rtol=1e-13)


def test_mvnun_matches_scalar_fast_path():
    rng = np.random.default_rng(1)
    for _ in range(200):
        s0, s1, r = rng.uniform(0.3, 2.0), rng.uniform(0.3, 2.0), rng.uniform(-0.8, 0.8)
        cov = np.array([[s0 * s0, r * s0 * s1], [r * s0 * s1, s1 * s1]])
        mean = rng.normal(0, 2, 2)
        lo = mean + rng.normal(0, 2, 2) - np.array([0.75, 0.8])
        hi = lo + np.array([1.5, 1.6])
        a, _ = bvn.mvnun(lo, hi, mean, cov)
        b, _ = bvn.mvnun_fast(lo, hi, mean, cov)
        assert abs(a - b) < 1e-15
