"""QSS simulator front end (mirror of spline_traj_optm/simulator)."""
