"""Candidate racing lines = smooth lateral-offset perturbations of a track's centre line.

The reference has no candidate generator; its optimisers move spline control points one at
a time and re-simulate (optimization/optimizer.py:163-341).  Here a candidate is a vector of lateral offsets
o_i (left positive, RaceTrack convention models/race_track.py:87-96) at the M centre samples, clipped to the
track bounds minus a safety margin; candidate 0 is the centre line itself.
"""
import numpy as np


def smooth_offsets(M, B, dist_left, dist_right, seed=1234, harmonics=8, amplitude=1.5, margin=1.0):
    """offsets[B, M] = clip(sum_h A_bh sin(h 2 pi i / M + phi_bh), -(right - margin), left - margin) with
    A_bh ~ N(0, amplitude / h), phi ~ U[0, 2 pi); row 0 is all zeros.  Deterministic in (seed, b)."""
    i = np.arange(M, dtype=np.float64)
    off = np.zeros((B, M), dtype=np.float64)
    lo = -np.maximum(np.asarray(dist_right, dtype=np.float64) - margin, 0.0)
    hi = np.maximum(np.asarray(dist_left, dtype=np.float64) - margin, 0.0)
    chunk = 4096
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        rng = np.random.default_rng([seed, b0])
        amp = rng.normal(0.0, 1.0, size=(nb, harmonics)) * (amplitude / np.arange(1, harmonics + 1))
        phi = rng.uniform(0.0, 2.0 * np.pi, size=(nb, harmonics))
        acc = np.zeros((nb, M))
        for h in range(1, harmonics + 1):
            acc += amp[:, h - 1:h] * np.sin(h * 2.0 * np.pi * i[None, :] / M + phi[:, h - 1:h])
        off[b0:b0 + nb] = np.clip(acc, lo[None, :], hi[None, :])
    off[0] = 0.0
    return off


def smooth_offsets_device(M, lo, hi, dist_left, dist_right, device, seed=1234, harmonics=8, amplitude=1.5, margin=1.0):
    """Candidates lo .. hi-1 of a (conceptually unbounded) batch, generated ON THE DEVICE and returned sample-major
    [M, round_up(hi - lo, 32)] (the layout the kernels read): the same family of lines as `smooth_offsets` - 8 sine
    harmonics, amplitudes ~ N(0, 1.5 / h) m, clipped to the bounds minus the margin, candidate 0 = the centre line -
    with torch's generator instead of NumPy's (different random lines; the big configurations only need them
    synthetic).  A candidate's offsets depend on (seed, its global index) only, never on how the batch is sharded."""
    import torch
    n = hi - lo
    ld = (n + 31) & ~31
    out = torch.zeros((M, ld), dtype=torch.float64, device=device)
    ang = (2.0 * np.pi / M) * torch.arange(M, dtype=torch.float64, device=device)
    lo_b = -torch.clamp(torch.as_tensor(np.asarray(dist_right, dtype=np.float64), device=device) - margin, min=0.0)
    hi_b = torch.clamp(torch.as_tensor(np.asarray(dist_left, dtype=np.float64), device=device) - margin, min=0.0)
    scale = amplitude / torch.arange(1, harmonics + 1, dtype=torch.float64, device=device)
    chunk = 4096
    for c in range(lo // chunk, (hi + chunk - 1) // chunk):
        g = torch.Generator(device=device)
        g.manual_seed(int(seed) * 1000003 + c)
        amp = torch.randn((chunk, harmonics), generator=g, dtype=torch.float64, device=device) * scale
        phi = torch.rand((chunk, harmonics), generator=g, dtype=torch.float64, device=device) * (2.0 * np.pi)
        b0, b1 = max(lo, c * chunk), min(hi, (c + 1) * chunk)
        rows = slice(b0 - c * chunk, b1 - c * chunk)
        acc = torch.zeros((M, b1 - b0), dtype=torch.float64, device=device)
        for h in range(1, harmonics + 1):
            acc += amp[rows, h - 1][None, :] * torch.sin(h * ang[:, None] + phi[rows, h - 1][None, :])
        out[:, b0 - lo:b1 - lo] = torch.minimum(torch.maximum(acc, lo_b[:, None]), hi_b[:, None])
    if lo == 0:
        out[:, 0] = 0.0
    return out


def banked_oval(straight=800.0, radius=250.0, width=15.0, bank_deg=9.0, blend=100.0, spacing=8.0):
    """Synthetic banked oval (BASELINE config 4): 4-column ``x, y, z, bank`` centre line plus left/right bounds.
    Two straights + two semicircles, counter-clockwise; bank ramps 0 -> bank_deg over ``blend`` metres into each
    curve.  Returns (centre[n, 4], left[n, 3], right[n, 3])."""
    per = 2.0 * straight + 2.0 * np.pi * radius
    n = int(per // spacing)
    s = np.arange(n) * (per / n)
    x, y, yaw, bank = (np.zeros(n) for _ in range(4))
    arc = np.pi * radius
    for k, sk in enumerate(s):
        if sk < straight:                                  # bottom straight, heading +x
            x[k], y[k], yaw[k] = -straight / 2 + sk, -radius, 0.0
            d = min(sk, straight - sk)
            bank[k] = bank_deg * max(0.0, 1.0 - d / blend)
        elif sk < straight + arc:                          # right curve
            a = (sk - straight) / radius
            x[k], y[k], yaw[k] = straight / 2 + radius * np.sin(a), -radius * np.cos(a), a
            bank[k] = bank_deg
        elif sk < 2 * straight + arc:                      # top straight, heading -x
            t = sk - straight - arc
            x[k], y[k], yaw[k] = straight / 2 - t, radius, np.pi
            d = min(t, straight - t)
            bank[k] = bank_deg * max(0.0, 1.0 - d / blend)
        else:                                              # left curve
            a = (sk - 2 * straight - arc) / radius
            x[k], y[k], yaw[k] = -straight / 2 - radius * np.sin(a), radius * np.cos(a), np.pi + a
            bank[k] = bank_deg
    nx, ny = -np.sin(yaw), np.cos(yaw)
    z = np.zeros(n)
    centre = np.stack([x, y, z, np.deg2rad(bank)], axis=1)
    left = np.stack([x + 0.5 * width * nx, y + 0.5 * width * ny, z], axis=1)
    right = np.stack([x - 0.5 * width * nx, y - 0.5 * width * ny, z], axis=1)
    return centre, left, right
