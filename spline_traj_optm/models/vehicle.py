from spline_trajectory_optimization_b200.models.vehicle import *  # noqa: F401,F403
from spline_trajectory_optimization_b200.models import vehicle as _impl
__all__ = [n for n in dir(_impl) if not n.startswith('_')]
