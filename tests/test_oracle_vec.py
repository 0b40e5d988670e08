"""The vectorised oracle (oracle/vec.py) against the reference's golden fixtures.  CPU only."""
import numpy as np
import pytest

from helpers import golden_cases, load_golden
from oracle import vec


@pytest.mark.parametrize("name", golden_cases())
def test_vec_reproduces_reference(name):
    step, ref = load_golden(name)
    out = vec.score_dense(step)
    keep = np.flatnonzero(out["level"] != vec.LEVEL_SKIPPED)
    C = len(ref["nT"])
    assert len(keep) == C
    np.testing.assert_array_equal(out["level"][keep], ref["level"])
    np.testing.assert_array_equal(out["reason"][keep], ref["reason"])
    np.testing.assert_array_equal(out["T"][keep], ref["nT"])
    for li, c in enumerate(keep):
        T = int(ref["nT"][li])
        np.testing.assert_allclose(out["traj"][:, c, :T], ref["traj"][:, li, :T], rtol=1e-9, atol=1e-9)
    if ref["prob"].size:
        got = out["prob"][keep].transpose(0, 2, 1)
        np.testing.assert_allclose(got[:, :, :ref["prob"].shape[2]], ref["prob"], rtol=1e-10, atol=1e-15)
        np.testing.assert_allclose(out["red"][:, keep], ref["red"], rtol=1e-10, atol=1e-15)
    hc = ref["has_cost"]
    np.testing.assert_allclose(out["cost"][12][keep][hc], ref["cost"][hc], rtol=1e-10, atol=1e-12)
    assert out["best"]["list_index"] == int(ref["best"])
    assert out["best"]["n_scored"] == C


def test_profile_subsets_agree_with_the_full_run():
    step, _ = load_golden("planning_15x15")
    full = vec.score_dense(step)
    part = vec.score_dense(step, profiles=[3, 7, 8])
    nd = len(step.d_list)
    for k, p in enumerate([3, 7, 8]):
        a = slice(k * nd, (k + 1) * nd)
        b = slice(p * nd, (p + 1) * nd)
        np.testing.assert_array_equal(part["level"][a], full["level"][b])
        np.testing.assert_array_equal(part["cost"][:, a], full["cost"][:, b])
        np.testing.assert_array_equal(part["prob"][a], full["prob"][b])
