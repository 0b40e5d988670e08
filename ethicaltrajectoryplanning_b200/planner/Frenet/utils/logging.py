"""Per-step log line of the Frenet planner: same header, separators and JSON conventions as the reference's
``planner/Frenet/utils/logging.py:10-61,166-182`` so ``analyze_tools`` / ``get_data_from_line`` read it unchanged."""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np

HEADER = "simulation_step;time;valid_trajectories;invalid_trajectories;predictions;calc_time_avg;reach_set"


def default(obj):
    """numpy -> JSON (logging.py:166-182)."""
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    raise TypeError("Not serializable (type: " + str(type(obj)) + ")")


class FrenetLogging:
    """Writes one text line per planning step (logging.py:31-61)."""

    def __init__(self, log_path: str) -> None:
        Path(os.path.dirname(log_path) or ".").mkdir(parents=True, exist_ok=True)
        self.__log_path = log_path
        with open(self.__log_path, "w+") as fh:
            fh.write(HEADER)

    def log_data(self, simulation_step, time, valid_trajectories, invalid_trajectories, predictions, calc_time_avg,
                 reach_set) -> None:
        with open(self.__log_path, "a") as fh:
            fh.write("\n" + ";".join(json.dumps(v, default=default) for v in (
                simulation_step, time, valid_trajectories, invalid_trajectories, predictions, calc_time_avg, reach_set)))


def get_data_from_line(file_path_in: str, line_num: int):
    """Parsed fields of one log line (the reader the reference's analysis tools use)."""
    with open(file_path_in) as fh:
        lines = fh.read().split("\n")
    keys = lines[0].split(";")
    vals = lines[line_num].split(";")
    return {k: json.loads(v) for k, v in zip(keys, vals)}
