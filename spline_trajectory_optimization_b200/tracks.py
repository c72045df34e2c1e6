"""Bundled track inputs.  ``data/monza_enu.npz`` holds the x, y columns of the reference's example CSVs
(spline_traj_optm/examples/race_track/monza/MONZA_{UNOPTIMIZED_LINE,LEFT_BOUNDARY,RIGHT_BOUNDARY}_enu.csv, loaded
with the reference's own loader arguments, tests/test_trajectory.py:12) so the benchmark input travels with the
package; regenerate with oracle/gen_golden.py."""
import os

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def monza_raw():
    d = np.load(os.path.join(DATA, "monza_enu.npz"))
    return d["center"], d["left"], d["right"]


def load_xy_csv(path):
    """x, y columns of a reference-style track CSV (header row, comma separated; trailing commas tolerated)."""
    rows = []
    with open(path) as f:
        next(f)
        for line in f:
            parts = [p for p in line.strip().split(",") if p != ""]
            if len(parts) >= 2:
                rows.append([float(v) for v in parts[:4]])
    width = min(len(r) for r in rows)
    return np.array([r[:width] for r in rows], dtype=np.float64)
