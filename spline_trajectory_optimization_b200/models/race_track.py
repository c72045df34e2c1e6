"""Race track = three closed splines + their discretisations + per-sample bounds.

Mirror of the reference's ``spline_traj_optm/models/race_track.py`` (RaceTrack.__init__ :9-85,
fill_trajectory_boundaries :98-104, frenet_to_global :87-96) for the parts the batched evaluator needs: the
splines (s=10, k=3 by default), their ``interval``-spaced samples, the left/right bound look-up and the lateral
coordinate convention (positive to the left).  The CasADi interpolants (:61-85) only serve the min-time NLP and
are out of scope; ``frenet_to_global`` is provided in NumPy on the sampled centre line instead.
"""
import numpy as np

from .trajectory import BSplineTrajectory, Trajectory


class RaceTrack:
    def __init__(self, name: str, left: np.ndarray, right: np.ndarray, centerline: np.ndarray, s=10.0, interval=2.0,
                 arc_length: bool = False, device: bool = True) -> None:
        """arc_length=True also fills DIST_TO_SF_* of the three discretisations with the reference's adaptive
        quadrature (seconds per line; the evaluator does not need it).  device=False keeps the whole construction on
        the host (SciPy evaluation, NumPy bound look-up): no CUDA library is loaded."""
        self._device = device
        assert left.shape[0] >= 3 and right.shape[0] >= 3
        assert left.shape[1] >= 2 and right.shape[1] >= 2
        self.name = name
        self.left_s = BSplineTrajectory(left[:, :2], s, 3)
        self.right_s = BSplineTrajectory(right[:, :2], s, 3)
        self.center_s = BSplineTrajectory(centerline[:, :2], s, 3)
        self.left_d = self.left_s.sample_along(interval, arc_length=arc_length, device=device)
        self.right_d = self.right_s.sample_along(interval, arc_length=arc_length, device=device)
        self.center_d = self.center_s.sample_along(interval, arc_length=arc_length, device=device)
        # closed boundary polylines (what the reference wraps in shapely LinearRings, :31-33)
        self.left_r = self.left_d[:, :2].copy()
        self.right_r = self.right_d[:, :2].copy()
        self.center_r = self.center_d[:, :2].copy()
        self.fill_trajectory_boundaries(self.center_d)
        c = self.center_d
        self.dist_to_left = np.linalg.norm(c[:, [Trajectory.X, Trajectory.Y]]
                                           - c[:, [Trajectory.LEFT_BOUND_X, Trajectory.LEFT_BOUND_Y]], axis=1)
        self.dist_to_right = np.linalg.norm(c[:, [Trajectory.X, Trajectory.Y]]
                                            - c[:, [Trajectory.RIGHT_BOUND_X, Trajectory.RIGHT_BOUND_Y]], axis=1)
        # optional per-centre-sample bank angle (rad); the reference never fills BANK
        if centerline.shape[1] >= 4:
            raw_bank = centerline[:, 3]
            # carried to the samples by parameter: nearest raw point in chord-length parameter
            seg = np.linalg.norm(np.diff(np.vstack([centerline[:, :2], centerline[:1, :2]]), axis=0), axis=1)
            u = np.concatenate([[0.0], np.cumsum(seg)]) / np.sum(seg)
            self.center_d[:, Trajectory.BANK] = np.interp(self.center_d.ts(), u, np.append(raw_bank, raw_bank[0]))

    def fill_trajectory_boundaries(self, traj: Trajectory):
        """Fills LEFT/RIGHT_BOUND_X/Y of ``traj`` in place (normal-ray look-up, max_dist = 100 m)."""
        traj.fill_bounds(self.left_r, self.right_r, max_dist=100.0, device=self._device)

    def left_normals(self):
        """Unit left normal (-sin yaw, cos yaw) at every centre sample."""
        yaw = self.center_d[:, Trajectory.YAW]
        return np.stack([-np.sin(yaw), np.cos(yaw)], axis=1)

    def frenet_to_global(self, idx, t, xi=0.0):
        """Centre sample index + lateral offset t (left positive) + heading offset -> (x, y, heading)."""
        c = self.center_d
        yaw = c[idx, Trajectory.YAW]
        return np.stack([c[idx, Trajectory.X] - np.sin(yaw) * t, c[idx, Trajectory.Y] + np.cos(yaw) * t,
                         np.arctan2(np.sin(yaw + xi), np.cos(yaw + xi))], axis=-1)
