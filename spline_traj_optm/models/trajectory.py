from spline_trajectory_optimization_b200.models.trajectory import *  # noqa: F401,F403
from spline_trajectory_optimization_b200.models import trajectory as _impl
__all__ = [n for n in dir(_impl) if not n.startswith('_')]
