"""Alias of spline_trajectory_optimization_b200.models."""
