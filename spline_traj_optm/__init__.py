"""Drop-in import name of the reference package: ``spline_traj_optm.models.*`` / ``.simulator.*`` resolve to the
B200-native implementation in ``spline_trajectory_optimization_b200``."""
