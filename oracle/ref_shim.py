"""ORACLE (test infrastructure, not product code): import the *unmodified* reference
modules from ``/root/reference`` under shims for its un-vendored dependencies.

Only usable inside the build container (the GPU box has no ``/root/reference``);
it exists to (1) generate the golden fixtures under ``tests/golden/`` with the
reference's own functions and (2) validate the NumPy restatements in
``oracle/port.py`` / ``oracle/vec.py``.  Nothing here is imported by the product.

Shim recipe:
  1. the code imports itself both as ``EthicalTrajectoryPlanning.x`` and ``x``
     (e.g. planner/Frenet/utils/frenet_functions.py:21-29) -> a temp dir with a
     symlink ``EthicalTrajectoryPlanning -> /root/reference`` on ``sys.path``;
  2. ``MagicMock`` modules for commonroad, commonroad_dc, commonroad_helper_functions,
     pygeos, shapely, matplotlib, prediction, agent_sim, progressbar, git;
  3. ``np.float = float`` (harm_estimation.py:248-249 needs numpy < 1.24);
  4. ``scipy.stats.mvn`` -> object whose ``mvnun`` is the four-corner Genz BVU
     (scipy removed the Fortran wrapper; ``scipy.special._ufuncs._bivariate_normal_cdf``
     is scipy's own port of the same BVU routine, upper-tail convention);
  5. ``pycrcc.RectOBB`` stand-in (center / r_x / local_x_axis);
  6. road boundary "never left": ``trajectories_collision_static_obstacles -> [-1]``,
     ``trajectory_preprocess_obb_sum -> (obj, 0)``;
  7. a scenario stub whose ``obstacle_by_id(id).obstacle_type`` returns attributes of the
     same mocked ``ObstacleType`` that harm_estimation imported.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import math
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np

REF_ROOT = os.environ.get("ETP_REFERENCE_ROOT", "/root/reference")

_MOCK_TOPLEVEL = (
    "commonroad",
    "commonroad_dc",
    "commonroad_helper_functions",
    "pygeos",
    "shapely",
    "matplotlib",
    "prediction",
    "agent_sim",
    "progressbar",
    "git",
    "networkx",
    "botorch",
    "statsmodels",
    "seaborn",
    "pandas_bokeh",
)


class _MockModule(types.ModuleType):
    """A module whose every attribute is a stable MagicMock (and which acts as a package)."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []  # package-like: allows ``import a.b.c``
        self._mocks = {}

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        m = self._mocks.get(item)
        if m is None:
            m = mock.MagicMock(name=f"{self.__name__}.{item}")
            self._mocks[item] = m
        return m


class _MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _MOCK_TOPLEVEL:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _MockModule(spec.name)

    def exec_module(self, module):
        return None


class _Mvn:
    """Stand-in for the removed ``scipy.stats.mvn`` (Fortran MVNDST wrapper).

    For N = 2 MVNDST evaluates BVNMVN = BVU(l1,l2) - BVU(u1,l2) - BVU(l1,u2) + BVU(u1,u2);
    scipy's ``_bivariate_normal_cdf`` ufunc is its C++ port of BVU.
    """

    @staticmethod
    def mvnun(lower, upper, means, covar):
        from scipy.special._ufuncs import _bivariate_normal_cdf as bvu

        covar = np.asarray(covar, dtype=float)
        s0 = math.sqrt(covar[0][0])
        s1 = math.sqrt(covar[1][1])
        r = covar[1][0] / s1 / s0
        xl = (lower[0] - means[0]) / s0
        xu = (upper[0] - means[0]) / s0
        yl = (lower[1] - means[1]) / s1
        yu = (upper[1] - means[1]) / s1
        p = bvu(xl, yl, r) - bvu(xu, yl, r) - bvu(xl, yu, r) + bvu(xu, yu, r)
        return float(p), 0


class _RectOBB:
    def __init__(self, r_x, r_y, orientation, x, y):
        self._rx, self._ry, self._o, self._c = r_x, r_y, orientation, np.array([x, y], dtype=float)

    def center(self):
        return self._c

    def r_x(self):
        return self._rx

    def r_y(self):
        return self._ry

    def local_x_axis(self):
        return np.array([math.cos(self._o), math.sin(self._o)])


class RefSpline:
    """``calc_position`` stand-in taking the same coefficient table the product takes."""

    def __init__(self, knots, coef_x, coef_y):
        self.s = np.asarray(knots, dtype=float)
        self.cx = np.asarray(coef_x, dtype=float)
        self.cy = np.asarray(coef_y, dtype=float)

    def calc_position(self, s):
        s = np.asarray(s, dtype=float)
        i = np.clip(np.searchsorted(self.s, s, side="right") - 1, 0, len(self.s) - 2)
        ds = s - self.s[i]
        cx, cy = self.cx[i], self.cy[i]
        x = cx[:, 0] + cx[:, 1] * ds + cx[:, 2] * ds * ds + cx[:, 3] * ds * ds * ds
        y = cy[:, 0] + cy[:, 1] * ds + cy[:, 2] * ds * ds + cy[:, 3] * ds * ds * ds
        return x, y


class _Namespace(types.SimpleNamespace):
    pass


_LOADED = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "risk_assessment"))


def load_reference():
    """Import the reference hot-path modules; returns a namespace of them (cached)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    link_dir = tempfile.mkdtemp(prefix="etp_refshim_")
    os.symlink(REF_ROOT, os.path.join(link_dir, "EthicalTrajectoryPlanning"))
    for p in (link_dir, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.meta_path.insert(0, _MockFinder())
    if not hasattr(np, "float"):
        np.float = float  # Q12
    if not hasattr(np, "VisibleDeprecationWarning"):
        np.VisibleDeprecationWarning = DeprecationWarning
    import scipy.stats

    scipy.stats.mvn = _Mvn()

    E = "EthicalTrajectoryPlanning."
    ns = _Namespace()
    ns.collision_probability = importlib.import_module(E + "risk_assessment.collision_probability")
    ns.harm_estimation = importlib.import_module(E + "risk_assessment.harm_estimation")
    ns.risk_costs = importlib.import_module(E + "risk_assessment.risk_costs")
    ns.validity_checks = importlib.import_module(E + "planner.Frenet.utils.validity_checks")
    ns.responsibility = importlib.import_module(E + "planner.utils.responsibility")
    ns.properties = importlib.import_module(E + "risk_assessment.helpers.properties")
    # frenet_functions imports its siblings un-prefixed (frenet_functions.py:21-28) and
    # risk_costs prefixed (29): import it un-prefixed like the reference's own __main__ does.
    ns.frenet_functions = importlib.import_module("planner.Frenet.utils.frenet_functions")
    ns.calc_trajectory_cost = importlib.import_module("planner.Frenet.utils.calc_trajectory_cost")
    ns.polynomials = importlib.import_module("planner.Frenet.utils.polynomials")
    ns.load_json = importlib.import_module("planner.Frenet.configs.load_json")
    ns.vehicleparams = importlib.import_module("planner.utils.vehicleparams")
    ns.logging = importlib.import_module("planner.Frenet.utils.logging")
    # un-prefixed twins that frenet_functions / calc_trajectory_cost actually call
    ns.validity_checks_unpref = importlib.import_module("planner.Frenet.utils.validity_checks")
    ns.risk_costs_unpref = importlib.import_module("risk_assessment.risk_costs")

    ns.collision_probability.mvn = _Mvn()
    ns.collision_probability.pycrcc.RectOBB = _RectOBB
    never_left = lambda **k: [-1]  # noqa: E731
    for mod in (ns.risk_costs, ns.risk_costs_unpref):
        mod.trajectory_queries.trajectories_collision_static_obstacles = never_left
    for name in list(sys.modules):
        if name.endswith("validity_checks"):
            m = sys.modules[name]
            m.trajectory_queries.trajectory_preprocess_obb_sum = lambda o: (o, 0)
            m.trajectory_queries.trajectories_collision_static_obstacles = never_left
            # collision with predictions is a pycrcc OBB sweep (un-vendored): host mask, all-pass
            m.collision_checker_prediction = lambda **k: False
    ns.ObstacleType = ns.harm_estimation.ObstacleType
    _LOADED = ns
    return ns


class ScenarioStub:
    """``scenario.obstacle_by_id(id).obstacle_type`` over a {id: type-name} table."""

    def __init__(self, ns, types_by_id):
        self._ns = ns
        self._types = dict(types_by_id)

    def obstacle_by_id(self, oid):
        name = self._types[oid]
        return types.SimpleNamespace(obstacle_type=getattr(self._ns.ObstacleType, name))


class EgoStateStub:
    def __init__(self, position, orientation, time_step=0, velocity=0.0, acceleration=0.0):
        self.position = np.asarray(position, dtype=float)
        self.orientation = float(orientation)
        self.time_step = int(time_step)
        self.velocity = float(velocity)
        self.acceleration = float(acceleration)
