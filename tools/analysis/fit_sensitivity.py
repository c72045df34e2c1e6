"""Developer tool (CPU): how sensitive is the reference's lap time to last-bit differences in the spline coefficients?

For bench candidates: lap through FITPACK's own coefficients (scipy splprep s=0, k=3, per=1: the reference's fit,
models/trajectory.py:219-220) -> the oracle's sampler + QSS, against the same chain started from (a) the one-lane Thomas
solve (oracle) and (b) the partitioned solve (host build of the device code).  All three coefficient sets agree to
~1e-14 relative; the QSS schedule is discontinuous in its inputs, so a few lines per thousand land on the other side of
a stop / overwrite decision and their laps move by 1e-6 .. 1e-4 s - for EVERY solver that is not FITPACK bit for bit.
Usage: python tools/analysis/fit_sensitivity.py [n_candidates]"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
import bench  # noqa: E402
import hostsim_py as H  # noqa: E402
import oracle_py as O  # noqa: E402
from helpers import golden, knots_from_u, veh_args  # noqa: E402
from scipy.interpolate import splprep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rt = bench.build_track()
off = bench.make_offsets(rt, 4096, 1234)[:n]
ctr, nrm, ts = rt.center_d[:, :2], rt.left_normals(), rt.center_d.ts()
pts = ctr[None] + off[:, :, None] * nrm[None]
ov = O.make_vehicle(*veh_args(golden("cand_m579_n579")))
zero = np.zeros(len(ts))


def lap_from(t, cx, cy, ref_pow):
    X, Y, YAW, R = O.sample(t, cx, cy, 3, ts, ref_pow)
    return O.qss(X, Y, R, zero, ov, ref_pow)["lap"]


hu, hcx, hcy, _ = H.fit_points(pts, -1)
rows = []
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for b in range(n):
        c = np.vstack([pts[b], pts[b][:1]])
        (t, (fx, fy), k), u = splprep([c[:, 0], c[:, 1]], s=0.0, k=3, per=1)
        ot, ocx, ocy = O.fit_periodic_cubic(pts[b])
        assert np.array_equal(t, ot) and np.array_equal(knots_from_u(hu[b]), t)
        ref = lap_from(t, fx, fy, 1)                      # the reference: FITPACK coefficients, Python pow
        rows.append((ref, lap_from(t, fx, fy, 0), lap_from(ot, ocx, ocy, 0), lap_from(t, hcx[b], hcy[b], 0),
                     np.max(np.abs(ocx - fx) / np.abs(fx).max()), np.max(np.abs(hcx[b] - fx) / np.abs(fx).max())))
r = np.array(rows)
for name, col in (("FITPACK coefficients, x*x instead of pow(x, 2)", 1), ("one-lane Thomas solve (oracle / mode 0)", 2),
                  ("partitioned solve (default)", 3)):
    d = np.abs(r[:, col] - r[:, 0])
    print(f"{name:48s} |dlap| vs reference: median {np.median(d):.1e}  95% {np.quantile(d, 0.95):.1e}  max {d.max():.1e}  "
          f"> 1e-6 s: {int((d > 1e-6).sum())} of {n}")
d = np.abs(r[:, 2] - r[:, 3])
print(f"{'Thomas vs partitioned':48s} |dlap|: median {np.median(d):.1e}  max {d.max():.1e}  > 1e-6 s: {int((d > 1e-6).sum())} of {n}")
print(f"max relative coefficient difference vs FITPACK: Thomas {r[:, 4].max():.1e}, partitioned {r[:, 5].max():.1e}")
