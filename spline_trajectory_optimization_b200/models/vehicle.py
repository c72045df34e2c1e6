"""Vehicle limits and lookup tables.

Mirror of the reference's ``spline_traj_optm/models/vehicle.py`` (VehicleParams :6-15, Vehicle :18-47): same
names, argument meaning and return values.  The two speed->acceleration tables are turned into piecewise
cubics by the same SciPy ``CubicSpline`` call as the reference, on the host; the resulting PPoly break points
and coefficients are what crosses the C ABI (``sto_vehicle_f64``), so the device evaluates the identical
polynomial.
"""
from dataclasses import dataclass

import numpy as np
from scipy.interpolate import CubicSpline


@dataclass
class VehicleParams:
    acc_speed_lookup: np.ndarray   # [n, 2] speed (m/s) -> max acceleration (m/s^2, positive)
    dcc_speed_lookup: np.ndarray   # [n, 2] speed (m/s) -> max deceleration (m/s^2, negative)
    max_lon_acc_mpss: float        # > 0
    max_lon_dcc_mpss: float        # < 0
    max_left_acc_mpss: float       # > 0
    max_right_acc_mpss: float      # < 0
    max_speed_mps: float
    max_jerk: float


class Vehicle:
    def __init__(self, param: VehicleParams):
        self.param = param
        acc = np.asarray(param.acc_speed_lookup, dtype=np.float64)
        dcc = np.asarray(param.dcc_speed_lookup, dtype=np.float64)
        self.acc_intp = CubicSpline(acc[:, 0], acc[:, 1])
        self.dcc_intp = CubicSpline(dcc[:, 0], dcc[:, 1])

    def lookup_acc_from_speed(self, speed_mps):
        return self.acc_intp(speed_mps)

    def lookup_dcc_from_speed(self, speed_mps):
        return self.dcc_intp(speed_mps)

    def lookup_acc_circle(self, lat=None, lon=None, model="ellipse"):
        """Friction-ellipse coupling: given one acceleration component, the (+, -) limits of the other."""
        assert (lat is not None) or (lon is not None)
        if model != "ellipse":
            return None
        p = self.param
        if lat is not None:
            val = np.clip(lat, p.max_right_acc_mpss, p.max_left_acc_mpss)
            semi = p.max_left_acc_mpss if val > 0.0 else p.max_right_acc_mpss
            pos, neg = p.max_lon_acc_mpss, p.max_lon_dcc_mpss
        else:
            val = np.clip(lon, p.max_lon_dcc_mpss, p.max_lon_acc_mpss)
            semi = p.max_lon_acc_mpss if val > 0.0 else p.max_lon_dcc_mpss
            pos, neg = p.max_left_acc_mpss, p.max_right_acc_mpss
        scale = np.sqrt(1.0 - val ** 2 / semi ** 2)
        return pos * scale, neg * scale
