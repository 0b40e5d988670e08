"""Golden vectors of the prediction enrichment, written by the UNMODIFIED reference functions
``get_orientation_velocity_and_shape_of_prediction`` (planner/Frenet/utils/prediction_helpers.py:75-135) and
``assign_responsibility_by_action_space`` (planner/utils/responsibility.py:57-89) under the shims of oracle/ref_shim.py.

    python tests/golden/make_golden_enrich.py        # build container only: /root/reference does not travel
"""
from __future__ import annotations

import copy
import importlib
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402


def scenes():
    rng = np.random.default_rng(2024)
    for case in range(6):
        n_obs = [5, 7, 3, 6, 8, 4][case]
        dt = [0.1, 0.1, 0.2, 0.1, 0.05, 0.1][case]
        raw, shapes, ori0, vel0 = {}, {}, {}, {}
        for i in range(n_obs):
            oid = 300 + 10 * case + i
            n = int(rng.integers(2, 30))
            kind = (i + case) % 6
            if kind == 0:
                n = 1                                     # one-sample prediction: initial orientation / velocity
            heading = rng.uniform(-np.pi, np.pi)
            speed = rng.uniform(0.5, 20.0)
            if kind == 1:
                speed = 0.0                               # static: "barely moves"
            if kind == 2:
                heading = rng.uniform(-np.pi + 0.1, -np.pi / 2 - 0.1)   # moving towards -x, -y: all dx, dy < 1e-4 (sic)
            k = np.arange(n, dtype=np.float64)
            p0 = rng.uniform(-200.0, 200.0, size=2)
            pos = p0[None] + (speed * dt * k)[:, None] * np.array([np.cos(heading), np.sin(heading)])[None]
            if kind == 3:
                pos = pos + rng.normal(0.0, 0.05, size=pos.shape)       # noisy positions
            if kind == 4:
                pos[:, 1] = p0[1]                                       # pure +x / -x motion: dy == 0 exactly
            raw[oid] = {"pos_list": pos, "cov_list": np.tile(np.eye(2) * 0.1, (n, 1, 1))}
            shapes[oid] = (float(rng.uniform(0.4, 12.0)), float(rng.uniform(0.4, 2.6)))
            ori0[oid] = float(rng.uniform(-4.0, 4.0))
            vel0[oid] = float(rng.uniform(0.0, 15.0))
        ego_pos = rng.uniform(-200.0, 200.0, size=2)
        ego_ori = [0.3, 3.0, -3.0, 1.5, -1.5, 3.1][case]               # near +-pi: the view cone is not wrapped (Q9)
        yield case, dt, raw, shapes, ori0, vel0, ego_pos, ego_ori


def main():
    ns = ref_shim.load_reference()
    ph = importlib.import_module("EthicalTrajectoryPlanning.planner.Frenet.utils.prediction_helpers")
    out = {}
    for case, dt, raw, shapes, ori0, vel0, ego_pos, ego_ori in scenes():
        class Scenario:
            pass

        sc = Scenario()
        sc.dt = dt
        sc.obstacle_by_id = lambda oid: SimpleNamespace(
            initial_state=SimpleNamespace(orientation=ori0[oid], velocity=vel0[oid]),
            obstacle_shape=SimpleNamespace(length=shapes[oid][0], width=shapes[oid][1]))
        pred = copy.deepcopy(raw)
        pred = ph.get_orientation_velocity_and_shape_of_prediction(predictions=pred, scenario=sc)
        ego = SimpleNamespace(position=ego_pos, orientation=ego_ori)
        pred = ns.responsibility.assign_responsibility_by_action_space(sc, ego, pred)
        oids = list(raw.keys())
        out[f"c{case}_ids"] = np.array(oids)
        out[f"c{case}_dt"] = dt
        out[f"c{case}_ego"] = np.array([ego_pos[0], ego_pos[1], ego_ori])
        for j, oid in enumerate(oids):
            out[f"c{case}_{j}_pos"] = raw[oid]["pos_list"]
            out[f"c{case}_{j}_in"] = np.array([shapes[oid][0], shapes[oid][1], ori0[oid], vel0[oid]])
            out[f"c{case}_{j}_yaw"] = np.asarray(pred[oid]["orientation_list"], dtype=np.float64)
            out[f"c{case}_{j}_vel"] = np.asarray(pred[oid]["v_list"], dtype=np.float64)
            out[f"c{case}_{j}_out"] = np.array([pred[oid]["shape"]["length"], pred[oid]["shape"]["width"],
                                                 pred[oid]["responsibility"]])
    out["n_cases"] = 6
    np.savez_compressed(os.path.join(HERE, "enrich_cases.npz"), **out)
    print("wrote", os.path.join(HERE, "enrich_cases.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
