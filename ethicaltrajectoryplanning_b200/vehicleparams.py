"""Ego vehicle constants consumed by the scoring path.

Attribute names follow the reference's ``planner/utils/vehicleparams.py:4-121``
(``l, w, l_r, m, steering.max, lateral_a_max, longitudinal.v_max, longitudinal.a_max``)
so a reference ``VehicleParameters`` object can be passed in unchanged.
"""
from __future__ import annotations

from types import SimpleNamespace

# name: (l, l_f*2.595, l_r*2.595, w, m, steer_max, v_min, v_max, v_switch)
_TABLE = {
    "ford_escort": (4.298, 2.9, 4.95, 1.674, 1050, 0.910, -13.9, 45.8, 4.755),
    "bmw_320i": (4.508, 3.793293, 4.667707, 1.610, 1475, 1.066, -13.6, 50.8, 7.319),
    "vw_vanagon": (4.569, 3.775563, 4.334437, 1.844, 1450, 1.023, -11.2, 41.7, 7.824),
}


class VehicleParameters:
    """Vehicle parameter set selected by name."""

    def __init__(self, vehicle_string="bmw_320i"):
        if vehicle_string not in _TABLE:
            raise ValueError("Value has to be ford_escort, bmw_320i or vw_vanagon")
        l, lf, lr, w, m, st, vmin, vmax, vsw = _TABLE[vehicle_string]
        self.l = l
        self.l_f = lf / 2.595
        self.l_r = lr / 2.595
        self.w = w
        self.m = m
        self.steering = SimpleNamespace(min=-st, max=st, v_min=-0.4, v_max=0.4)
        self.longitudinal = SimpleNamespace(v_min=vmin, v_max=vmax, v_switch=vsw, a_max=11.5)
        self.lateral_a_max = 10.0
