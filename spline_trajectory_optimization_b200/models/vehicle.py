"""Vehicle limits and speed-dependent acceleration tables (host side).

API mirror of the reference's ``spline_traj_optm/models/vehicle.py`` (``VehicleParams`` :6-15, ``Vehicle`` :18-47) so
that callers written against the reference keep working: same constructor fields, same method names, same return
shapes.  What differs is what the object is *for*: here it is the host-side description that gets flattened into the
``sto_vehicle_f64`` POD (``evaluator.vehicle_struct``) and evaluated on the GPU by ``ppoly4`` / ``max_lat_acc`` in
``csrc/sto_common.cuh``.  The two tables are interpolated by SciPy's ``CubicSpline`` exactly as in the reference, and
the resulting piecewise-polynomial break points and coefficients cross the C ABI unchanged, so device and reference
evaluate the identical polynomial.
"""
from dataclasses import dataclass

import numpy as np
from scipy.interpolate import CubicSpline

_G_AXES = {
    # component given -> (lower clip, upper clip, semi-axis if positive, semi-axis otherwise, + limit, - limit)
    "lon": ("max_lon_dcc_mpss", "max_lon_acc_mpss", "max_lon_acc_mpss", "max_lon_dcc_mpss",
            "max_left_acc_mpss", "max_right_acc_mpss"),
    "lat": ("max_right_acc_mpss", "max_left_acc_mpss", "max_left_acc_mpss", "max_right_acc_mpss",
            "max_lon_acc_mpss", "max_lon_dcc_mpss"),
}


@dataclass
class VehicleParams:
    """Field order and names are the reference's (positional construction is common in its scripts)."""
    acc_speed_lookup: np.ndarray   # rows (speed m/s, max acceleration m/s^2 > 0)
    dcc_speed_lookup: np.ndarray   # rows (speed m/s, max deceleration m/s^2 < 0)
    max_lon_acc_mpss: float        # friction-ellipse semi-axis, accelerating (> 0)
    max_lon_dcc_mpss: float        # friction-ellipse semi-axis, braking (< 0)
    max_left_acc_mpss: float       # lateral semi-axis, left (> 0)
    max_right_acc_mpss: float      # lateral semi-axis, right (< 0); the QSS never reads it
    max_speed_mps: float
    max_jerk: float


def _table(rows):
    rows = np.asarray(rows, dtype=np.float64)
    return CubicSpline(rows[:, 0], rows[:, 1])


class Vehicle:
    def __init__(self, param: VehicleParams):
        self.param = param
        self.acc_intp = _table(param.acc_speed_lookup)   # PPoly: .x break points, .c[4][n-1] coefficients
        self.dcc_intp = _table(param.dcc_speed_lookup)

    def lookup_acc_from_speed(self, speed_mps):
        """Largest acceleration available at this speed (0-d array, as SciPy returns it)."""
        return self.acc_intp(speed_mps)

    def lookup_dcc_from_speed(self, speed_mps):
        """Largest deceleration (negative) available at this speed."""
        return self.dcc_intp(speed_mps)

    def _ellipse(self, axis, value):
        lo, hi, semi_pos, semi_neg, lim_pos, lim_neg = (getattr(self.param, name) for name in _G_AXES[axis])
        value = np.clip(value, lo, hi)
        semi = semi_pos if value > 0.0 else semi_neg
        remaining = np.sqrt(1.0 - value ** 2 / semi ** 2)   # same expression order as the reference (:47)
        return lim_pos * remaining, lim_neg * remaining

    def lookup_acc_circle(self, lat=None, lon=None, model="ellipse"):
        """Friction-ellipse coupling: with one acceleration component given (``lat`` wins if both are), the
        (positive, negative) limits of the other one.  Only the 'ellipse' model exists, as in the reference."""
        assert (lat is not None) or (lon is not None)
        if model != "ellipse":
            return None
        return self._ellipse("lat", lat) if lat is not None else self._ellipse("lon", lon)
