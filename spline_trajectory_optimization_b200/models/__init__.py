"""Host-side mirrors of the reference's data classes (spline_traj_optm/models)."""
