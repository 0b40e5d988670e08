#!/usr/bin/env python
"""Parity report of the CUDA path: worst deviations from the golden fixtures of the unmodified reference, from the
vectorised fp64 oracle on random scenes and on the full 10^4-candidate step.  Writes markdown to stdout."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # the oracle is test infrastructure: this report lives with the tests

import numpy as np  # noqa: E402


def relerr(got, ref, floor):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + floor))) if got.size else 0.0


def main():
    from helpers import golden_cases, load_golden

    from ethicaltrajectoryplanning_b200 import _abi
    from ethicaltrajectoryplanning_b200.scoring import ScoringContext
    from ethicaltrajectoryplanning_b200.synthetic import make_step
    from oracle import vec

    spec = importlib.util.spec_from_file_location("rs", os.path.join(ROOT, "tests", "test_gpu_random_scenes.py"))
    rs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rs)

    print("# Parity report (round 1)\n")
    print("Relative errors are |got - ref| / (|ref| + floor) with floor 1e-5 for probabilities / risks (the absolute floor of "
          "the parity bar is 1e-9), 1e-12 for harm, 1e-9 for costs; trajectories against max(|ref|, 1e-3).\n")
    for precision in ("fp32", "fp64"):
        ctx = ScoringContext(0, precision)
        worst = dict(prob=0.0, risk=0.0, harm=0.0, cost=0.0, traj=0.0)
        n_cases = n_cand = mism_level = mism_best = 0
        for name in golden_cases():
            step, ref = load_golden(name)
            res = ctx.plan(step)
            out = res.numpy()
            keep = np.flatnonzero(out["level"] != _abi.LEVEL_SKIPPED)
            n_cases += 1
            n_cand += len(keep)
            mism_level += int(np.sum(out["level"][keep].astype(np.int64) != ref["level"]))
            mism_level += int(np.sum(out["reason"][keep].astype(np.int64) != ref["reason"]))
            mism_best += int(res.best["list_index"] != int(ref["best"]))
            if ref["prob"].size:
                got = out["prob"][keep].transpose(0, 2, 1)[:, :, :ref["prob"].shape[2]]
                worst["prob"] = max(worst["prob"], relerr(got, ref["prob"], 1e-5))
                worst["risk"] = max(worst["risk"], relerr(out["red"][:2, keep], ref["red"][:2], 1e-5))
                worst["harm"] = max(worst["harm"], relerr(out["red"][2:, keep], ref["red"][2:], 1e-12))
            hc = ref["has_cost"]
            total = out["cost"][_abi.COST_NAMES.index("total")][keep]
            worst["cost"] = max(worst["cost"], relerr(total[hc], ref["cost"][hc], 1e-9))
            for li, c in enumerate(keep):
                T = int(ref["nT"][li])
                want = ref["traj"][:, li, :T]
                worst["traj"] = max(worst["traj"], float(np.max(np.abs(out["traj"][:, c, :T] - want) / np.maximum(np.abs(want), 1e-3))))
        print(f"## {precision}: golden fixtures of the unmodified reference ({n_cases} cases, {n_cand} candidates)\n")
        print(f"* validity level / reason mismatches: **{mism_level}**; selected-trajectory mismatches: **{mism_best}**")
        print("* worst relative error: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()) + "\n")
        worst = dict(prob=0.0, risk=0.0, harm=0.0, cost=0.0)
        n_scene = n_cand = mism = best_mism = 0
        for seed in range(60):
            step = rs.random_scene(seed)
            ref = vec.score_dense(step)
            res = ctx.plan(step)
            out = res.numpy()
            n_scene += 1
            n_cand += int(np.sum(ref["level"] != vec.LEVEL_SKIPPED))
            mism += int(np.sum(out["level"].astype(np.int64) != ref["level"])) + int(np.sum(out["reason"].astype(np.int64) != ref["reason"]))
            best_mism += int(res.best["index"] != ref["best"]["index"])
            worst["prob"] = max(worst["prob"], relerr(out["prob"], ref["prob"], 1e-5))
            worst["risk"] = max(worst["risk"], relerr(out["red"][:2], ref["red"][:2], 1e-5))
            worst["harm"] = max(worst["harm"], relerr(out["red"][2:], ref["red"][2:], 1e-12))
            live = ref["level"] != vec.LEVEL_SKIPPED
            worst["cost"] = max(worst["cost"], relerr(out["cost"][12, live], ref["cost"][12, live], 1e-9))
        print(f"## {precision}: {n_scene} random scenes against the vectorised fp64 oracle ({n_cand} candidates; paths in every "
              "direction, |rho| up to 0.9, all harm models)\n")
        print(f"* validity level / reason mismatches: **{mism}**; selected-trajectory mismatches: **{best_mism}**")
        print("* worst relative error: " + ", ".join(f"{k} {v:.2e}" for k, v in worst.items()) + "\n")
        step = make_step("cfg2")
        ref = vec.score_dense(step)
        res = ctx.plan(step)
        out = res.numpy()
        print(f"## {precision}: cfg2, the full 10^4-candidate step\n")
        print(f"* level mismatches: **{int(np.sum(out['level'].astype(np.int64) != ref['level']))}**, selected index "
              f"{res.best['index']} vs {ref['best']['index']}")
        print(f"* worst relative error: prob {relerr(out['prob'], ref['prob'], 1e-5):.2e}, risk "
              f"{relerr(out['red'][:2], ref['red'][:2], 1e-5):.2e}, harm {relerr(out['red'][2:], ref['red'][2:], 1e-12):.2e}, "
              f"cost {relerr(out['cost'][12], ref['cost'][12], 1e-9):.2e}\n")
        ctx.close()


if __name__ == "__main__":
    main()
