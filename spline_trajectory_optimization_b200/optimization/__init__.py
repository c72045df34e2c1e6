"""Batched front end for the callers of the hot path (reference spline_traj_optm/optimization)."""
