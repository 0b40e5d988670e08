"""Drop-in for ``planner/Frenet/frenet_planner.py``: ``FrenetPlanner`` with the reference's constructor arguments
(:79-92) and planning step (``_step_planner``, :265-576).

What stays as in the reference: the ego Frenet state bookkeeping (:293-299), the end-velocity rule (:301-319), the
order calc_frenet_trajectories -> predictions -> sort_frenet_trajectories -> sort by cost -> best -> log line ->
``self._trajectory`` (:326-340, 433-457, 469-477, 555-576).  What changes: the two calls are fused CUDA launches.

Everything CommonRoad-specific that the reference builds in its constructor (global path + reference spline
``planning.py:321-377``, WaleNet predictor, road boundary, collision checker, goal area) is injected through
keyword arguments instead, because those packages are not part of the scored path.
"""
from __future__ import annotations

import copy
from types import SimpleNamespace

import numpy as np

from ... import dropin
from ...configs import load_harm_parameter_json, load_risk_json, load_weight_json
from ...synthetic import PlanningStep, assign_responsibility_by_action_space, planner_v_list
from ..utils.timers import ExecTimer, device_stages
from .utils.frenet_functions import calc_frenet_trajectories, sort_frenet_trajectories
from .utils.logging import ColumnarLogging, FrenetLogging
from .utils.validity_checks import VALIDITY_LEVELS  # noqa: F401

DEFAULT_FRENET_PARAMETERS = {  # frenet_planner.py:116-128 (planning_fast.json values)
    "t_list": [2.0], "v_list_generation_mode": "linspace", "n_v_samples": 5,
    "d_list": np.linspace(-3.5, 3.5, 15), "dt": 0.1, "v_thr": 3.0,
}


class FrenetPlanner:
    """Jerk optimal planning in frenet coordinates with quintic / quartic polynomials, scored on the GPU."""

    min_trajectory_length = 2

    def __init__(self, scenario, planning_problem, ego_id: int, vehicle_params, mode, exec_timer=None,
                 frenet_parameters: dict = None, sensor_radius: float = 50.0, plot_frenet_trajectories: bool = False,
                 weights=None, settings=None, *, reference_spline=None, global_path=None, predictor=None, ego_state=None,
                 road_boundary=None, collision_checker=None, goal_area=None, log_path=None, obstacle_types=None,
                 materialize=True, precision=None, columnar_log_path=None):
        if mode not in ("ground_truth", "WaleNet", "risk"):
            raise ValueError("mode must be ground_truth, WaleNet, or risk")
        if reference_spline is None:
            raise ValueError("reference_spline (the CubicSpline2D of planning.py:375-377) must be given")
        self.scenario, self.planning_problem, self.ego_id = scenario, planning_problem, ego_id
        # frenet_planner.py:94-97: a disabled timer stands in when none is given
        self.exec_timer = ExecTimer(timing_enabled=False) if exec_timer is None else exec_timer
        self.frenet_parameters = dict(DEFAULT_FRENET_PARAMETERS if frenet_parameters is None else frenet_parameters)
        self.p = vehicle_params
        self.params_harm = load_harm_parameter_json()
        self.params_weights = load_weight_json() if weights is None else weights
        self.params_mode = settings["risk_dict"] if (settings is not None and "risk_dict" in settings) else load_risk_json()
        self.params_dict = {"weights": self.params_weights, "modes": self.params_mode, "harm": self.params_harm}
        if self.params_weights.get("responsibility", 0) > 0:
            raise NotImplementedError("weights['responsibility'] > 0 needs reachable sets (frenet_planner.py:205-215): "
                                      "outside the scored path")
        goal0 = planning_problem.goal.state_list[0] if hasattr(planning_problem, "goal") else None
        if goal0 is not None and hasattr(goal0, "velocity"):
            self.v_goal_min, self.v_goal_max = goal0.velocity.start, goal0.velocity.end
        else:
            self.v_goal_min = self.v_goal_max = None
        self.sensor_radius, self.mode = sensor_radius, mode
        self.plot_frenet_trajectories = plot_frenet_trajectories
        self.predictor = predictor
        self.reference_spline, self.global_path = reference_spline, global_path
        self.road_boundary, self.collision_checker, self.goal_area = road_boundary, collision_checker, goal_area
        self.reach_set, self.responsibility = None, False
        self.obstacle_types = obstacle_types
        self.logger = FrenetLogging(log_path) if log_path else None
        # binary side-car of the per-step log for the fused path (logging.ColumnarLogging; legacy line on demand)
        self.columnar_logger = ColumnarLogging(columnar_log_path) if columnar_log_path else None
        self.materialize = materialize
        self.ctx = dropin.default_context(precision)
        self.ego_state = ego_state
        if self.ego_state is not None and not hasattr(self.ego_state, "acceleration"):
            self.ego_state.acceleration = 0.0  # frenet_planner.py:166-167
        self.time_step = 0 if ego_state is None else getattr(ego_state, "time_step", 0)
        n = self.min_trajectory_length
        self._trajectory = {k: np.zeros(n) for k in ("s_loc_m", "d_loc_m", "d_d_loc_mps", "d_dd_loc_mps2", "x_m", "y_m",
                                                     "psi_rad", "kappa_radpm", "v_mps", "ax_mps2")}
        self._trajectory["time_s"] = np.arange(0, 0.1 * n, 0.1)  # planning.py:251-263
        self.last_best = None

    @property
    def trajectory(self):
        return self._trajectory

    # -- planning.py:265-300 ---------------------------------------------------------------------------------
    def step(self, scenario, current_lanelet_id, time_step, ego_state, prediction=None, v_max=50):
        self.scenario, self.time_step, self.ego_state = scenario, time_step, ego_state
        self._prediction = prediction
        self._step_planner()

    # -- frenet_planner.py:342-421 ----------------------------------------------------------------------------
    def _predictions(self, t_list):
        if self.mode not in ("WaleNet", "risk"):
            return None
        if self.predictor is None:
            return {}
        if hasattr(self.predictor, "step"):
            raw = self.predictor.step(time_step=self.ego_state.time_step, obstacle_id_list=None, scenario=self.scenario)
        else:
            raw = self.predictor(self.ego_state.time_step)
        predictions = copy.copy(raw)
        self._resident = False
        self._raw_record = None
        if predictions and "orientation_list" not in next(iter(predictions.values())):
            # get_orientation_velocity_and_shape_of_prediction (prediction_helpers.py:75-135) and
            # assign_responsibility_by_action_space (responsibility.py:57-89) on the device, straight into the packed tile
            # shape, initial state and class of an obstacle do not change during a scenario: looked up once per id
            static = self.__dict__.setdefault("_obstacle_static", {})
            prev = self.__dict__.get("_static_dicts")
            if prev is not None and prev[0] == tuple(predictions):
                # the same obstacles as in the previous cycle: the very same description dicts (the context recognises them)
                _, shapes, ori0, vel0, types = prev
            else:
                shapes, ori0, vel0, types = {}, {}, {}, {}
                for oid in predictions:
                    st = static.get(oid)
                    if st is None:
                        ob = self.scenario.obstacle_by_id(oid)
                        st = static[oid] = ((ob.obstacle_shape.length, ob.obstacle_shape.width), ob.initial_state.orientation,
                                            ob.initial_state.velocity,
                                            dropin.obstacle_type_names(self.scenario, {oid: None}, self.obstacle_types)[oid])
                    shapes[oid], ori0[oid], vel0[oid], types[oid] = st
                self._static_dicts = (tuple(predictions), shapes, ori0, vel0, types)
            self._resident_types = types
            self._resident_static = (shapes, ori0, vel0)  # what the enrichment on the device was given (tests replay it on the host)
            # fused path without a log: the record is only prepared here and travels with the step (ScoringContext.plan_cycle,
            # which also makes sure of the parameters)
            cycle = not self.materialize and self.columnar_logger is None
            if not cycle:
                self.ctx.ensure_params(self.params_dict, self.p, self.mode)
            self._raw_record = self.ctx.set_raw_predictions(predictions, shapes, ori0, vel0, types, self.scenario.dt,
                                                            np.asarray(self.ego_state.position), self.ego_state.orientation,
                                                            upload=not cycle)
            if not cycle:
                self._raw_record = None
            self._resident = True
            if self.materialize:  # the log line prints the enriched dict
                self.ctx.enriched(predictions)
            return predictions
        # assign_responsibility_by_action_space (responsibility.py:57-89)
        return assign_responsibility_by_action_space(np.asarray(self.ego_state.position), self.ego_state.orientation,
                                                     predictions)

    def _step_planner(self):
        """Frenet Planner step function (frenet_planner.py:265-576); timer labels as at :270-561."""
        self.exec_timer.start_timer("simulation/total")
        try:
            self._step_planner_timed()
        finally:
            self.exec_timer.stop_timer("simulation/total")

    def _step_planner_timed(self):
        fpar = self.frenet_parameters
        # find position along the reference spline (s, s_d, s_dd, d, d_d, d_dd)  (:293-299)
        c_s = self.trajectory["s_loc_m"][1]
        c_s_d = self.ego_state.velocity
        c_s_dd = self.ego_state.acceleration
        c_d = self.trajectory["d_loc_m"][1]
        c_d_d = self.trajectory["d_d_loc_mps"][1]
        c_d_dd = self.trajectory["d_dd_loc_mps2"][1]
        d_list, t_list = fpar["d_list"], fpar["t_list"]
        if fpar.get("v_list_generation_mode", "linspace") != "linspace":
            raise NotImplementedError("only the linspace v-list mode is part of the scored path (SURVEY.md 8a row a2)")
        with self.exec_timer.time_with_cm("simulation/get v list"):
            v_list = planner_v_list(c_s_d, t_list, self.p, fpar["n_v_samples"], self.v_goal_min, self.v_goal_max)  # :301-319
        with self.exec_timer.time_with_cm("simulation/prediction"):
            predictions = self._predictions(t_list)

        if not self.materialize:
            return self._step_fused(c_s, c_s_d, c_s_dd, c_d, c_d_d, c_d_dd, d_list, t_list, v_list, predictions)

        with self.exec_timer.time_with_cm("simulation/calculate trajectories/total"):  # :321
            ft_list = calc_frenet_trajectories(c_s=c_s, c_s_d=c_s_d, c_s_dd=c_s_dd, c_d=c_d, c_d_d=c_d_d, c_d_dd=c_d_dd,
                                               d_list=d_list, t_list=t_list, v_list=v_list, dt=fpar["dt"],
                                               csp=self.reference_spline, v_thr=fpar["v_thr"], exec_timer=self.exec_timer)
        with self.exec_timer.time_with_cm("simulation/sort trajectories/total"):  # :430
            ft_list_valid, ft_list_invalid, validity_dict = sort_frenet_trajectories(
                ego_state=self.ego_state, fp_list=ft_list, global_path=self.global_path, predictions=predictions,
                mode=self.mode, params=self.params_dict, planning_problem=self.planning_problem, scenario=self.scenario,
                vehicle_params=self.p, ego_id=self.ego_id, dt=fpar["dt"], sensor_radius=self.sensor_radius,
                road_boundary=self.road_boundary, collision_checker=self.collision_checker, goal_area=self.goal_area,
                exec_timer=self.exec_timer, reach_set=None, obstacle_types=self.obstacle_types)
            with self.exec_timer.time_with_cm("simulation/sort trajectories/sort list by costs"):  # :453-457
                ft_list_valid.sort(key=lambda fp: fp.cost, reverse=False)  # stable: ties keep generation order
        if self.logger is not None:  # :468-477
            with self.exec_timer.time_with_cm("log trajectories"):
                self.logger.log_data(self.time_step, self.time_step * fpar["dt"], [d.__dict__ for d in ft_list_valid],
                                     [d.__dict__ for d in ft_list_invalid], predictions, 0, None)
        best = ft_list_valid[0] if len(ft_list_valid) > 0 else ft_list_invalid[0]  # :555-558
        self.validity_dict = validity_dict
        self.last_best = best
        self._trajectory = {  # :564-576
            "s_loc_m": best.s, "d_loc_m": best.d, "d_d_loc_mps": best.d_d, "d_dd_loc_mps2": best.d_dd, "x_m": best.x,
            "y_m": best.y, "psi_rad": best.yaw, "kappa_radpm": best.curv, "v_mps": best.s_d, "ax_mps2": best.s_dd,
            "time_s": best.t,
        }

    def _step_fused(self, c_s, c_s_d, c_s_dd, c_d, c_d_d, c_d_dd, d_list, t_list, v_list, predictions):
        """Same selection without materialising per-candidate Python objects: the whole cycle is one library call
        (``ScoringContext.plan_cycle`` = ``etp_plan_cycle``: this cycle's predictions and step in one upload, three launches
        replayed as a CUDA graph, winner and its 13 fp64 rows written into pinned host memory); with a columnar log the
        materialised step."""
        fpar = self.frenet_parameters
        from .utils.calc_trajectory_cost import cost_context, goal_context
        from .utils.validity_checks import device_road_boundary

        vel_mode, vel_target, factor = cost_context(None, self.ego_state, self.planning_problem, self.scenario, fpar["dt"],
                                                    self.goal_area)
        if getattr(self, "_resident", False):
            types = self._resident_types
        else:
            types = dropin.obstacle_type_names(self.scenario, predictions, self.obstacle_types) if predictions else {}
        step = PlanningStep(
            c_s=float(c_s), c_s_d=float(c_s_d), c_s_dd=float(c_s_dd), c_d=float(c_d), c_d_d=float(c_d_d),
            c_d_dd=float(c_d_dd), v_list=np.asarray(v_list, dtype=np.float64), t_list=np.asarray(t_list, dtype=np.float64),
            d_list=np.asarray(d_list, dtype=np.float64), dt=float(fpar["dt"]), v_thr=float(fpar["v_thr"]),
            csp=self.reference_spline, predictions=predictions, obstacle_types=types,
            ego_position=np.asarray(self.ego_state.position, dtype=np.float64), ego_orientation=float(self.ego_state.orientation),
            vehicle=self.p, params=self.params_dict, mode=self.mode, velocity_cost_mode=vel_mode,
            velocity_target=vel_target, cost_factor=float(np.ndim(factor) == 0 and factor or 1.0),
            predictions_resident=getattr(self, "_resident", False),
            goal=goal_context(self.ego_state, self.planning_problem, self.scenario, self.goal_area),
            road_boundary=device_road_boundary(self.road_boundary))
        self.last_step = step
        # the fused step IS "calculate trajectories" + "sort trajectories": both labels of the reference are kept, the
        # kernels' own times go under "simulation/sort trajectories/device/..."
        getattr(self.exec_timer, "record", lambda *a: None)("simulation/calculate trajectories/total", 0.0)
        with self.exec_timer.time_with_cm("simulation/sort trajectories/total"), \
                device_stages(self.ctx, self.exec_timer, "simulation/sort trajectories/"):
            if self.columnar_logger is not None:
                # logging needs every candidate: materialised step, tensors appended as they come off the device
                res = self.ctx.plan(step)
                with self.exec_timer.time_with_cm("log trajectories"):
                    self.columnar_logger.log_step(self.time_step, self.time_step * fpar["dt"], res, step, predictions, 0, None)
                self.last_best, self._trajectory = res.best, res.best_trajectory()
                return
            if getattr(self, "_raw_record", None) is not None and getattr(self, "_resident", False):
                self.last_best, self._trajectory = self.ctx.plan_cycle(step, raw=self._raw_record)
            elif predictions is not None and not getattr(self, "_resident", False):
                self.last_best, self._trajectory = self.ctx.plan_cycle(step)
            else:
                self.last_best, self._trajectory = self.ctx.plan_winner(step)

    # -- planning.py:117-131 -----------------------------------------------------------------------------------
    def ideal_tracking_state(self):
        """Next ego state under ideal tracking: the 2nd entry of the planned trajectory (quirk Q7: v := s_d)."""
        tr = self._trajectory
        return SimpleNamespace(position=np.array([tr["x_m"][1], tr["y_m"][1]]), orientation=float(tr["psi_rad"][1]),
                               velocity=float(tr["v_mps"][1]), acceleration=float(tr["ax_mps2"][1]),
                               time_step=self.time_step + 1)
