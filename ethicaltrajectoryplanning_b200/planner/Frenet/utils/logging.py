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


# ------------------------------------------------------------------------------------------------------------
# Columnar side-car.  The reference's text line holds every candidate's __dict__ as JSON
# (>= 100 MB per scenario, README.md:82) and costs far more than scoring the candidates.  The side-car appends
# one binary record per planning step -- the result tensors as they come off the device, plus a small JSON header --
# and rebuilds the exact legacy line on demand (``legacy_line``), so analysis tools keep reading what they read today.
#
#   record = b"ETPL" | u64 header bytes | header JSON (utf-8) | arrays, 64-byte aligned, in header["arrays"] order
# ------------------------------------------------------------------------------------------------------------
MAGIC = b"ETPL"
_ARRAYS = ("traj", "red", "cost", "level", "reason", "bd_harm", "dense_index", "n_samples")


class ColumnarLogging:
    """Binary per-step log of a ``ScoreResult`` (one file per scenario, records appended)."""

    def __init__(self, log_path: str) -> None:
        Path(os.path.dirname(log_path) or ".").mkdir(parents=True, exist_ok=True)
        self.__log_path = log_path
        open(self.__log_path, "wb").close()

    def log_step(self, simulation_step, time, result, step, predictions, calc_time_avg=0, reach_set=None) -> int:
        """Append one record; ``result`` is the materialised ScoreResult of ``step``.  Returns the bytes written."""
        out = result.numpy()
        keep = np.flatnonzero(out["level"] != 255)  # generation order without the skipped slots = the reference's fp_list
        nt, nd = len(step.t_list), len(step.d_list)
        arrays = {
            "traj": np.ascontiguousarray(out["traj"][:, keep]), "red": np.ascontiguousarray(out["red"][:, keep]),
            "cost": np.ascontiguousarray(out["cost"][:, keep]), "level": np.ascontiguousarray(out["level"][keep]),
            "reason": np.ascontiguousarray(out["reason"][keep]), "bd_harm": np.ascontiguousarray(out["bd_harm"][keep]),
            "dense_index": keep.astype(np.int64),
            "n_samples": np.asarray(result.n_samples, dtype=np.int32)[(keep // nd) % nt],
        }
        header = {
            "simulation_step": simulation_step, "time": time, "calc_time_avg": calc_time_avg, "reach_set": reach_set,
            "obstacle_ids": list(result.obstacle_ids), "mode": step.mode, "dt": step.dt,
            "t_list": [float(t) for t in step.t_list], "nt": nt, "nd": nd,
            "weights": step.params["weights"], "multiple_cost_functions": bool(step.params["modes"]["multiple_cost_functions"]),
            "host_bd_harm": step.bd_harm is not None or getattr(step, "road_boundary", None) is not None,
            "predictions": json.loads(json.dumps(predictions, default=default)),
            "arrays": [[k, str(arrays[k].dtype), list(arrays[k].shape)] for k in _ARRAYS],
        }
        hb = json.dumps(header).encode("utf-8")
        n = 0
        with open(self.__log_path, "ab") as fh:
            n += fh.write(MAGIC + np.uint64(len(hb)).tobytes() + hb)
            for k in _ARRAYS:
                pad = (-fh.tell()) % 64
                n += fh.write(b"\0" * pad)
                n += fh.write(arrays[k].tobytes())
        return n


def read_columnar(file_path_in: str):
    """All records of a side-car file: list of (header dict, {name: ndarray})."""
    records = []
    with open(file_path_in, "rb") as fh:
        blob = fh.read()
    pos = 0
    while pos < len(blob):
        if blob[pos:pos + 4] != MAGIC:
            raise ValueError("not a columnar log record at byte %d" % pos)
        hl = int(np.frombuffer(blob, dtype=np.uint64, count=1, offset=pos + 4)[0])
        header = json.loads(blob[pos + 12:pos + 12 + hl].decode("utf-8"))
        pos += 12 + hl
        arrays = {}
        for name, dtype, shape in header["arrays"]:
            pos += (-pos) % 64
            cnt = int(np.prod(shape))
            arrays[name] = np.frombuffer(blob, dtype=np.dtype(dtype), count=cnt, offset=pos).reshape(shape)
            pos += cnt * np.dtype(dtype).itemsize
        records.append((header, arrays))
    return records


def legacy_line(header, arrays) -> str:
    """The text line ``FrenetLogging.log_data`` writes for this step (without the leading newline): every candidate's
    ``__dict__`` in the reference's key order, valid list cost-sorted, invalid list by level then generation order."""
    from .frenet_functions import FrenetTrajectory, fill_trajectories

    predictions = {int(k): v for k, v in header["predictions"].items()} if header["predictions"] is not None else None
    params = {"weights": header["weights"], "modes": {"multiple_cost_functions": header["multiple_cost_functions"]}}
    nt, nd = header["nt"], header["nd"]
    traj = arrays["traj"].astype(np.float64)
    fps = []
    for i, c in enumerate(arrays["dense_index"]):
        T = int(arrays["n_samples"][i])
        f = traj[:, i, :T]
        fps.append(FrenetTrajectory(t=np.arange(0.0, header["t_list"][(int(c) // nd) % nt], header["dt"])[:T], d=f[0], d_d=f[1],
                                    d_dd=f[2], d_ddd=f[3], s=f[4], s_d=f[5], s_dd=f[6], s_ddd=f[7], x=f[8], y=f[9], yaw=f[10],
                                    v=f[11], curv=f[12]))
    valid, invalid, _ = fill_trajectories(
        fps, arrays["red"].astype(np.float64), arrays["cost"].astype(np.float64), arrays["level"].astype(np.int64),
        arrays["reason"].astype(np.int64), header["obstacle_ids"], predictions,
        arrays["bd_harm"] if header["host_bd_harm"] else None, params, header["mode"])
    valid.sort(key=lambda fp: fp.cost, reverse=False)
    return ";".join(json.dumps(v, default=default) for v in (
        header["simulation_step"], header["time"], [d.__dict__ for d in valid], [d.__dict__ for d in invalid],
        header["predictions"] if predictions is None else predictions, header["calc_time_avg"], header["reach_set"]))
