"""Ego vehicle parameters (reference: planners/common/vehicle/vehicle.py:5-46).

Same constructor and attribute names.  ``polygon`` is the closed footprint ring as a ``[5, 2]``
float64 array instead of a shapely ``Polygon``: the collision check that consumed it
(frenet_optimal_planner.py:179) runs on the GPU from ``l`` and ``w``; nothing on this path
needs GEOS.
"""
import numpy as np


class Vehicle(object):
    def __init__(self, vehicle_params, safety_factor: float = 1.0):
        self.l: float = vehicle_params.l * safety_factor
        self.w: float = vehicle_params.w * safety_factor
        self.h: float = 1.5 * safety_factor
        self.bbox_size: np.ndarray = np.array([self.l, self.w, self.h])

        hl, hw = self.l / 2, self.w / 2
        # clockwise from front-left, closed (vehicle.py:24-30)
        self.corners = [(hl, hw), (hl, -hw), (-hl, -hw), (-hl, hw), (hl, hw)]
        self.polygon = np.array(self.corners, dtype=np.float64)

        self.a = vehicle_params.a
        self.b = vehicle_params.b
        self.L = self.a + self.b
        self.T_f = vehicle_params.T_f
        self.T_r = vehicle_params.T_r
        self.max_speed = vehicle_params.longitudinal.v_max
        self.max_accel = vehicle_params.longitudinal.a_max
        self.deccel = -self.max_speed / 5.0
        self.max_steering_angle = vehicle_params.steering.max
        self.max_steering_rate = vehicle_params.steering.v_max
        self.max_curvature = np.sin(self.max_steering_angle) / self.L
        self.max_kappa_d = vehicle_params.steering.kappa_dot_max
        self.max_kappa_dd = vehicle_params.steering.kappa_dot_dot_max
