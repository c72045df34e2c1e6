"""Trajectory table, closed B-spline line and TTL I/O.

Mirror of the reference's ``spline_traj_optm/models/trajectory.py``: ``Trajectory`` (:25-209), ``BSplineTrajectory``
(:212-309), ``save_ttl``/``load_ttl`` (:312-358), ``Region``/``Bound`` (:11-22) with the same names, column
numbers, argument meaning and quirks.  Differences are behind the API:

* ``BSplineTrajectory.sample_along`` evaluates position, heading and turn radius on the GPU
  (``sto_sample_spline_f64``, SciPy's ``_deBoor_D`` operation order) when a CUDA device is present;
* ``fill_bounds`` / ``fill_region`` are NumPy restatements (shapely is not required);
* the smoothing fit itself (``splprep`` with s > 0) and the adaptive arc-length quadrature stay on the host with
  the same SciPy calls as the reference; batches of candidate lines (s = 0, k = 3) are
  fitted on the GPU by ``BatchedLineEvaluator``.
"""
import copy
import ctypes as C
import pickle
import warnings
from dataclasses import dataclass

import numpy as np
from scipy import interpolate
from scipy.integrate import IntegrationWarning, quad
from scipy.interpolate import BSpline


@dataclass
class Region:
    name: str
    code: int
    vertices: np.ndarray  # [n, 2]


@dataclass
class Bound:
    name: str
    type: str
    vertices: np.ndarray


def _ray_ring_hits(points, normals, ring, max_dist):
    """Nearest intersection of the two-sided segment P -/+ max_dist * n with a closed polyline.

    points, normals: [n, 2]; ring: [m, 2] vertices (closed implicitly).  Returns hit[n, 2] and found[n].
    Restates Trajectory.fill_bounds' shapely query (reference models/trajectory.py:83-141): the whole 2*max_dist
    segment is intersected, the hit closest to P wins, no hit -> the point itself.
    """
    a = np.asarray(ring, dtype=np.float64)
    if np.array_equal(a[0], a[-1]):
        a = a[:-1]
    e = np.roll(a, -1, axis=0) - a                       # [m, 2] edge vectors
    n_pts = points.shape[0]
    hit = points.copy()
    found = np.zeros(n_pts, dtype=bool)
    step = max(1, int(4_000_000 // max(1, a.shape[0])))
    for lo in range(0, n_pts, step):
        P = points[lo:lo + step, None, :]                # [c, 1, 2]
        d = normals[lo:lo + step, None, :]
        w = a[None, :, :] - P                            # [c, m, 2]
        den = d[..., 0] * e[None, :, 1] - d[..., 1] * e[None, :, 0]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (w[..., 0] * e[None, :, 1] - w[..., 1] * e[None, :, 0]) / den   # along the normal
            s = (w[..., 0] * d[..., 1] - w[..., 1] * d[..., 0]) / den           # along the edge
        ok = (den != 0.0) & (s >= 0.0) & (s <= 1.0) & (np.abs(t) <= max_dist)
        tt = np.where(ok, np.abs(t), np.inf)
        j = np.argmin(tt, axis=1)
        rows = np.arange(tt.shape[0])
        has = np.isfinite(tt[rows, j])
        tsel = np.where(has, t[rows, j], 0.0)
        hp = points[lo:lo + step] + tsel[:, None] * normals[lo:lo + step]
        hit[lo:lo + step][has] = hp[has]
        found[lo:lo + step] = has
    return hit, found


def _device_ray_ring_hits(points, normals, ring, max_dist):
    """GPU version of _ray_ring_hits (sto_fill_bounds_f64); None when no CUDA device is visible."""
    try:
        import torch
    except ImportError:  # pragma: no cover
        return None
    if not torch.cuda.is_available():
        return None
    from .. import _lib
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    a = np.asarray(ring, dtype=np.float64)
    if np.array_equal(a[0], a[-1]):
        a = a[:-1]
    up = lambda x: torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    px, py, nx, ny, rg = up(points[:, 0]), up(points[:, 1]), up(normals[:, 0]), up(normals[:, 1]), up(a)
    n = len(points)
    out = torch.empty((2, n), dtype=torch.float64, device=dev)
    found = torch.empty(n, dtype=torch.int32, device=dev)
    _lib.check(lib.sto_fill_bounds_f64(px.data_ptr(), py.data_ptr(), nx.data_ptr(), ny.data_ptr(), n, rg.data_ptr(),
                                       len(a), float(max_dist), out[0].data_ptr(), out[1].data_ptr(), found.data_ptr(),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.T.cpu().numpy(), found.cpu().numpy().astype(bool)


class Trajectory:
    X = 0
    Y = 1
    Z = 2
    YAW = 3
    SPEED = 4
    CURVATURE = 5          # holds the turn RADIUS in metres (reference quirk, models/trajectory.py:253-260,281)
    DIST_TO_SF_BWD = 6
    DIST_TO_SF_FWD = 7
    REGION = 8
    LEFT_BOUND_X = 9
    LEFT_BOUND_Y = 10
    RIGHT_BOUND_X = 11
    RIGHT_BOUND_Y = 12
    BANK = 13
    LON_ACC = 14
    LAT_ACC = 15
    TIME = 16
    IDX = 17
    ITERATION_FLAG = 18
    NUM_COLUMNS = 19

    def __init__(self, num_point: int, ttl_num: int = 0, origin=None) -> None:
        self.ttl_num = ttl_num
        self.origin = origin
        self.points = np.zeros((num_point, Trajectory.NUM_COLUMNS), dtype=np.float64)
        self.points[:, Trajectory.IDX] = np.arange(num_point)
        self.points[:, Trajectory.ITERATION_FLAG] = -1

    def __getitem__(self, key):
        return self.points[key]

    def __setitem__(self, key, val):
        self.points[key] = val

    def __len__(self):
        return len(self.points)

    def __iter__(self):
        return iter(self.points)

    def copy(self):
        # like the reference, only the table is carried over (ttl_num / origin are dropped)
        dup = Trajectory(len(self.points))
        dup.points = self.points.copy()
        return dup

    def inc(self, idx: int):
        return 0 if idx + 1 == len(self.points) else idx + 1

    def dec(self, idx: int):
        return len(self.points) - 1 if idx - 1 < 0 else idx - 1

    def distance(self, pt1, pt2):
        return np.linalg.norm(pt1[Trajectory.X:Trajectory.Y + 1] - pt2[Trajectory.X:Trajectory.Y + 1])

    def ts(self):
        return np.linspace(0.0, 1.0, len(self), endpoint=False)

    # ---- bounds / regions (NumPy restatement of the shapely queries) ------------------------------------------
    def fill_bounds(self, left_poly, right_poly, max_dist=100.0, device=True):
        """left_poly / right_poly: [m, 2] vertex arrays of the closed boundary polylines (anything with a
        ``coords`` attribute, e.g. a shapely LinearRing, is accepted too).  Runs on the GPU when one is visible
        (same arithmetic as the NumPy path, bit-identical results); device=False forces NumPy."""
        def verts(poly):
            return np.asarray(poly.coords if hasattr(poly, "coords") else poly, dtype=np.float64)[:, :2]
        P = self.points[:, [Trajectory.X, Trajectory.Y]]
        yaw = self.points[:, Trajectory.YAW]
        for poly, norm, cx, cy in ((left_poly, np.pi / 2.0, Trajectory.LEFT_BOUND_X, Trajectory.LEFT_BOUND_Y),
                                   (right_poly, -np.pi / 2.0, Trajectory.RIGHT_BOUND_X, Trajectory.RIGHT_BOUND_Y)):
            n = np.stack([np.cos(yaw + norm), np.sin(yaw + norm)], axis=1)
            res = _device_ray_ring_hits(P, n, verts(poly), max_dist) if device else None
            hit, _ = res if res is not None else _ray_ring_hits(P, n, verts(poly), max_dist)
            self.points[:, cx] = hit[:, 0]
            self.points[:, cy] = hit[:, 1]

    def fill_region(self, regions: list):
        x, y = self.points[:, Trajectory.X], self.points[:, Trajectory.Y]
        assigned = np.zeros(len(self), dtype=bool)
        for region in regions:   # first containing polygon wins, as in the reference's early return
            v = np.asarray(region.vertices, dtype=np.float64)
            x0, y0 = v[:, 0], v[:, 1]
            x1, y1 = np.roll(x0, -1), np.roll(y0, -1)
            with np.errstate(divide="ignore", invalid="ignore"):
                cross = ((y0[None] > y[:, None]) != (y1[None] > y[:, None])) & (
                    x[:, None] < (x1 - x0)[None] * (y[:, None] - y0[None]) / (y1 - y0)[None] + x0[None])
            inside = (np.count_nonzero(cross, axis=1) % 2 == 1) & ~assigned
            self.points[inside, Trajectory.REGION] = region.code
            assigned |= inside

    # ---- derived columns ---------------------------------------------------------------------------------------
    def fill_distance(self):
        nxt = np.roll(self.points[:, :2], -1, axis=0)
        seg = np.linalg.norm(self.points[:, :2] - nxt, axis=1)
        bwd = np.concatenate([[0.0], seg[:-1]])
        self.points[:, Trajectory.DIST_TO_SF_BWD] = np.cumsum(bwd)
        self.points[:, Trajectory.DIST_TO_SF_FWD] = np.sum(seg) - self.points[:, Trajectory.DIST_TO_SF_BWD]

    def fill_time(self):
        """Per-segment travel time into TIME[next] (NOT accumulated; TIME[0] = closing segment), as the
        reference does (models/trajectory.py:158-180).  The lap time is ``TIME.sum()``."""
        if np.any(self.points[:, Trajectory.SPEED] == 0.0):   # the reference's guard (:160-163), without its typo
            raise Exception("Zero speed and lon_acc encoutered. Cannot fill time.")
        self.points[0, Trajectory.TIME] = 0.0
        n = len(self.points)
        for this in range(n):
            nxt = this + 1 if this + 1 < n else 0
            x = self.distance(self.points[this], self.points[nxt])
            self.points[nxt, Trajectory.TIME] = x / (0.5 * (self.points[this, Trajectory.SPEED]
                                                           + self.points[nxt, Trajectory.SPEED]))

    def lap_time(self):
        return float(np.sum(self.points[:, Trajectory.TIME]))

    def save(f, traj):
        np.savetxt(f, traj.points, delimiter=",")

    def load(f):
        arr = np.loadtxt(f, np.float64, delimiter=",")
        traj = Trajectory(len(arr))
        traj.points = arr
        return traj


def _device_sample(t, cx, cy, k, ts):
    """x, y, yaw, radius at ts on the GPU (sto_sample_spline_f64); None when no CUDA device is visible."""
    try:
        import torch
    except ImportError:  # pragma: no cover
        return None
    if not torch.cuda.is_available():
        return None
    from .. import _lib
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    d = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev) for a in (t, cx, cy, ts)]
    n = len(ts)
    out = torch.empty((4, n), dtype=torch.float64, device=dev)
    _lib.check(lib.sto_sample_spline_f64(d[0].data_ptr(), len(t), d[1].data_ptr(), d[2].data_ptr(), int(k),
                                         d[3].data_ptr(), n, out[0].data_ptr(), out[1].data_ptr(),
                                         out[2].data_ptr(), out[3].data_ptr(),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy()


def _device_arc_sections(t, cx, cy, k, ts):
    """Lengths of the curve between consecutive samples on the GPU (sto_arc_sections_f64)."""
    import torch
    from .. import _lib
    if not torch.cuda.is_available():
        raise RuntimeError('arc_length="gauss" needs a CUDA device')
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    d = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev) for a in (t, cx, cy, ts)]
    sec = torch.empty(len(ts), dtype=torch.float64, device=dev)
    _lib.check(lib.sto_arc_sections_f64(d[0].data_ptr(), len(t), d[1].data_ptr(), d[2].data_ptr(), int(k),
                                        d[3].data_ptr(), len(ts), sec.data_ptr(),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return sec.cpu().numpy()


class BSplineTrajectory:
    def __init__(self, coordinates: np.ndarray, s: float, k: int):
        assert coordinates.shape[0] >= 3 and coordinates.shape[1] == 2 and len(
            coordinates.shape) == 2, "coordinates should be N * 2"
        closed = np.vstack([coordinates, coordinates[:1]])
        tck, _ = interpolate.splprep([closed[:, 0], closed[:, 1]], s=s, per=True, k=k)
        self._spl_x = BSpline(tck[0], tck[1][0], tck[2])
        self._spl_y = BSpline(tck[0], tck[1][1], tck[2])
        self._length = self._section_length(0.0, 1.0)

    # speed |r'(t)| and its integral
    def _speed(self, t):
        return np.sqrt(interpolate.splev(t, self._spl_x, der=1) ** 2 + interpolate.splev(t, self._spl_y, der=1) ** 2)

    def _section_length(self, t_min, t_max):
        # same call as the reference (models/trajectory.py:228-230); its whole-lap integral always trips QUADPACK's
        # round-off IntegrationWarning, which is noise here
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", IntegrationWarning)
            length, _ = quad(self._speed, t_min, t_max, limit=1000)
        return length

    def eval_sectional_length(self, ts):
        return self._section_length(ts[0], ts[1])

    def _eval_d_sectional_length(self, spl, ts):
        length, _ = quad(lambda t: interpolate.splev(t, spl, der=2) / self._speed(t), ts[0], ts[1], limit=200)
        return length

    def eval_dx_sectional_length(self, ts):
        return self._eval_d_sectional_length(self._spl_x, ts)

    def eval_dy_sectional_length(self, ts):
        return self._eval_d_sectional_length(self._spl_y, ts)

    def eval(self, t, der=0):
        return interpolate.splev(t, self._spl_x, der=der), interpolate.splev(t, self._spl_y, der=der)

    def eval_yaw(self, t):
        return np.arctan2(interpolate.splev(t, self._spl_y, der=1), interpolate.splev(t, self._spl_x, der=1))

    def eval_turn_radius(self, t):
        dx, dy = self.eval(t, 1)
        d2x, d2y = self.eval(t, 2)
        curvature = np.abs(dx * d2y - dy * d2x) / np.sqrt((dx ** 2 + dy ** 2) ** 3)
        return 1.0 / np.abs(curvature)

    def get_length(self):
        return self._length

    def sample_along(self, interval: float = None, ts=None, arc_length=True, device: bool = True) -> Trajectory:
        """Uniform-in-parameter resampling (NOT uniform in arc length), as the reference.

        arc_length: True = DIST_TO_SF_BWD/FWD by the reference's per-sample adaptive quadrature on the host (>99 %
        of the reference's time in this call; the QSS never reads those columns); "gauss" = the same columns from
        the GPU's fixed-order Gauss-Legendre kernel (agrees with quad to ~1e-10 relative, ~1000x faster); False =
        leave them zero.  device=False forces the SciPy evaluation path for X/Y/YAW/CURVATURE.
        """
        if interval is not None:
            num_sample = int(self.get_length() // interval)
            ts = np.linspace(0.0, 1.0, num_sample, endpoint=False)
        ts = np.asarray(ts, dtype=np.float64)
        traj = Trajectory(len(ts))
        cols = _device_sample(self._spl_x.t, self._spl_x.c, self._spl_y.c, self._spl_x.k, ts) if device else None
        if cols is None:
            cols = (interpolate.splev(ts, self._spl_x), interpolate.splev(ts, self._spl_y), self.eval_yaw(ts),
                    self.eval_turn_radius(ts))
        traj[:, Trajectory.X], traj[:, Trajectory.Y] = cols[0], cols[1]
        traj[:, Trajectory.YAW], traj[:, Trajectory.CURVATURE] = cols[2], cols[3]
        if isinstance(arc_length, str) and arc_length == "gauss":
            sec = _device_arc_sections(self._spl_x.t, self._spl_x.c, self._spl_y.c, self._spl_x.k, ts)
            traj[:, Trajectory.DIST_TO_SF_BWD] = np.cumsum(sec)
            traj[:, Trajectory.DIST_TO_SF_FWD] = self._length - traj[:, Trajectory.DIST_TO_SF_BWD]
        elif arc_length:
            acc = 0.0
            for i in range(1, len(traj)):
                acc = acc + self._section_length(ts[i - 1], ts[i])
                traj[i, Trajectory.DIST_TO_SF_BWD] = acc
            traj[:, Trajectory.DIST_TO_SF_FWD] = self._length - traj[:, Trajectory.DIST_TO_SF_BWD]
        return traj

    def copy(self):
        return copy.deepcopy(self)

    def set_control_point(self, idx, coord):
        self._spl_x.c[idx] = coord[0]
        self._spl_y.c[idx] = coord[1]

    def get_control_point(self, idx):
        return self._spl_x.c[idx], self._spl_y.c[idx]

    def save(f, traj):
        with open(f, "wb") as fh:
            pickle.dump(traj, fh)

    def load(f):
        with open(f, "rb") as fh:
            return pickle.load(fh)


_TTL_COLUMNS = (Trajectory.X, Trajectory.Y, Trajectory.Z, Trajectory.YAW, Trajectory.SPEED, Trajectory.CURVATURE,
                Trajectory.DIST_TO_SF_BWD, Trajectory.DIST_TO_SF_FWD, Trajectory.REGION, Trajectory.LEFT_BOUND_X,
                Trajectory.LEFT_BOUND_Y, Trajectory.RIGHT_BOUND_X, Trajectory.RIGHT_BOUND_Y, Trajectory.BANK,
                Trajectory.LON_ACC, Trajectory.LAT_ACC, Trajectory.TIME)


def save_ttl(ttl_path: str, trajectory: Trajectory):
    """TTL wire format: header ``ttl_num,N,track_length[,origin x3]`` then 17 columns per row (REGION as int)."""
    head = [str(trajectory.ttl_num), str(len(trajectory)), str(trajectory[0, Trajectory.DIST_TO_SF_FWD])]
    if trajectory.origin is not None:
        head += [str(x) for x in trajectory.origin]
    with open(ttl_path, "w") as f:
        f.write(",".join(head) + "\n")
        for row in trajectory.points:
            f.write(",".join(str(int(row[c])) if c == Trajectory.REGION else str(row[c]) for c in _TTL_COLUMNS) + "\n")


def load_ttl(ttl_path: str) -> Trajectory:
    with open(ttl_path, "r") as f:
        header = f.readline().split(",")
    assert len(header) >= 6
    data = np.loadtxt(ttl_path, dtype=float, delimiter=",", skiprows=1)
    trajectory = Trajectory(len(data), int(header[0]), (float(header[3]), float(header[4]), float(header[5])))
    trajectory.points[:, :data.shape[1]] = data
    return trajectory
