"""The oracle (oracle/port.py, oracle/bvn.py) against the golden fixtures produced by the
UNMODIFIED reference functions (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from helpers import golden_cases, load_golden
from oracle import bvn, port

FAST = ["cfg1_ethical", "cfg1_ego", "cfg1_standard", "lowspeed_0p5", "lowspeed_2p9", "fast_30", "vmax_clamped",
        "multi_t", "short_pred", "long_pred", "no_obstacles", "max_risk_level3", "harm_lr1s", "harm_lr12s",
        "harm_lr12a", "harm_lr4a", "mahalanobis", "multi_cost_fn", "jerk_travelled", "velmode1_factor",
        "planning_15x15",
        # random scenes: paths in every direction (00, 03, 06, 09, 12 along the +-pi cut: un-wrapped impact angles, Q4)
        "scene_00", "scene_01", "scene_02", "scene_03", "scene_05", "scene_06", "scene_09", "scene_12"]


def test_every_fixture_is_listed():
    assert sorted(FAST) == golden_cases()


@pytest.mark.parametrize("name", FAST)
def test_port_reproduces_reference(name):
    step, ref = load_golden(name)
    best, valid, invalid, allfp = port.plan_step(step, mvnun=bvn.mvnun_fast)
    C = len(allfp)
    assert C == len(ref["nT"])
    oids = list(step.predictions.keys())
    for i, fp in enumerate(allfp):
        T = len(fp.t)
        assert T == ref["nT"][i]
        for f, nm in enumerate(port.Traj.FIELDS):
            np.testing.assert_array_equal(getattr(fp, nm), ref["traj"][f, i, :T], err_msg=f"{nm}[{i}]")
        assert fp.valid_level == ref["level"][i]
        assert port.REASONS.index(fp.reason_invalid) == ref["reason"][i]
        for j, o in enumerate(oids):
            p = np.asarray(fp.prob_dict[o], dtype=float)
            np.testing.assert_allclose(p, ref["prob"][i, j, :len(p)], rtol=1e-12, atol=1e-16)
            got = [fp.ego_risk_dict[o], fp.obst_risk_dict[o], fp.ego_harm_dict[o], fp.obst_harm_dict[o]]
            np.testing.assert_allclose(got, ref["red"][:, i, j], rtol=1e-12, atol=1e-16)
        np.testing.assert_allclose(fp.cost, ref["cost"][i], rtol=1e-12, atol=1e-14)
    assert allfp.index(best) == int(ref["best"])
    assert [allfp.index(f) for f in valid] == ref["valid_order"].tolist()
    assert [allfp.index(f) for f in invalid] == ref["invalid_order"].tolist()


def test_own_bvu_port_on_one_fixture():
    """Same fixture through the NumPy restatement of Genz' BVU instead of scipy's compiled one."""
    step, ref = load_golden("short_pred")
    best, valid, invalid, allfp = port.plan_step(step, mvnun=bvn.mvnun)
    oids = list(step.predictions.keys())
    for i, fp in enumerate(allfp):
        for j, o in enumerate(oids):
            p = np.asarray(fp.prob_dict[o], dtype=float)
            np.testing.assert_allclose(p, ref["prob"][i, j, :len(p)], rtol=1e-12, atol=1e-16)
    assert allfp.index(best) == int(ref["best"])


def test_empty_candidate_list_raises_like_the_reference():
    """Q16: everything skipped by ds <= |dT| -> max([]) ValueError (frenet_functions.py:562-564)."""
    step, _ = load_golden("lowspeed_0p5")
    step.d_list = np.array([50.0, -60.0])
    with pytest.raises(ValueError):
        port.plan_step(step, mvnun=bvn.mvnun_fast)
