"""Golden log lines written by the reference's own ``FrenetLogging.log_data`` (planner/Frenet/utils/logging.py:31-61,
166-182, imported unmodified) for the trajectory objects the reference's own ``sort_frenet_trajectories`` leaves behind
on committed fixture steps -- the call ``FrenetPlanner._step_planner`` makes (frenet_planner.py:469-477):

    logger.log_data(time_step, time_step * dt, [d.__dict__ for d in valid], [d.__dict__ for d in invalid], predictions, 0, None)

Run in the build container only:   python tests/golden/make_golden_logline.py
Writes tests/golden/logline_<case>.txt.gz (header line + the step's line).
"""
from __future__ import annotations

import gzip
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from helpers import load_golden  # noqa: E402
from make_golden import run_reference  # noqa: E402
from oracle import ref_shim  # noqa: E402

CASES = ("cfg1_ethical", "max_risk_level3", "multi_t", "no_obstacles")
TIME_STEP = 3


def main():
    ns = ref_shim.load_reference()
    for name in CASES:
        step, _ = load_golden(name)
        fp_list, valid, invalid, best, _ = run_reference(step)
        for fp in fp_list:
            del fp._gen_index  # bookkeeping of run_reference, not an attribute the reference sets
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "logs", "frenet.csv")
            logger = ns.logging.FrenetLogging(path)
            logger.log_data(TIME_STEP, TIME_STEP * step.dt, [d.__dict__ for d in valid], [d.__dict__ for d in invalid],
                            step.predictions, 0, None)
            text = open(path).read()
        out = os.path.join(HERE, f"logline_{name}.txt.gz")
        with gzip.GzipFile(out, "wb", mtime=0) as fh:
            fh.write(text.encode("utf-8"))
        print(f"{name}: {len(valid)} valid, {len(invalid)} invalid, {len(text)} characters -> {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
