"""Tabulated 6-DoF solid-body motion: the `constant/6DoF.dat` contract.

Format written by the reference (circularSloshingTank/generate_motion.py:13-42):
    N
    (
    (t (x y z) (rx ry rz))
    ...
    )
and consumed by OpenFOAM's `solidBodyMotionFunction sixDoFMotion` with `type table`
Function1s for translation (columns 0 1) and rotation (columns 0 2)
(circularSloshingTank/constant/dynamicMeshDict:17-44): rows are interpolated linearly and
clamped outside the table; rotation is in degrees, applied as the quaternion sequence XYZ
about CofG, then translated [OF13-MEM].
"""
from __future__ import annotations

import math
import re

import numpy as np

from .foamfile import FoamError


def smootherstep(tau):
    """6 tau^5 - 15 tau^4 + 10 tau^3 (generate_motion.py:5-7)."""
    return tau * tau * tau * (tau * (tau * 6 - 15) + 10)


def orbital_table(r_max, f, duration, dt, ramp_duration):
    """Rows of the orbital shaking table, restating generate_motion.py:8-42 (including its
    %.6g rounding, which is what OpenFOAM reads back).  ramp_duration < 0 means 10 % of the
    duration (generate_motion.py:57-63)."""
    if ramp_duration < 0:
        ramp_duration = duration * 0.1
    n_steps = int(duration / dt) + 1
    rows = np.zeros((n_steps, 7))
    for i in range(n_steps):
        ti = i * dt
        r = r_max * smootherstep(ti / ramp_duration) if ti < ramp_duration else r_max
        th = 2 * math.pi * f * ti
        rows[i] = [float(f"{v:.6g}") for v in (ti, r * math.cos(th), r * math.sin(th), 0.0, 0.0, 0.0, 0.0)]
    return rows


def gen6dof_table(n_times=100, end_time=40.0):
    """Restatement of sloshingTank3D6DoF/gen6DoF/gen6DoF.C:44-82: translation amplitude
    (2 3 2) m at (0.5 0.8 0.4) rad/s, rotation amplitude (30 10 10) deg at (0.4 0.7 0.5)
    rad/s, all sine, sampled at n_times points over [0, end_time]."""
    t = np.linspace(0.0, end_time, n_times)  # gen6DoF.C:49-52 (t = i*endTime/(nTimes-1))
    ta, tw = np.array([2.0, 3.0, 2.0]), np.array([0.5, 0.8, 0.4])
    ra, rw = np.array([30.0, 10.0, 10.0]), np.array([0.4, 0.7, 0.5])
    rows = np.zeros((n_times, 7))
    rows[:, 0] = t
    rows[:, 1:4] = ta[None, :] * np.sin(tw[None, :] * t[:, None])
    rows[:, 4:7] = ra[None, :] * np.sin(rw[None, :] * t[:, None])
    return rows


def write_table(path, rows, fmt="%.6g"):
    with open(path, "w") as f:
        f.write(f"{rows.shape[0]}\n(\n")
        for r in rows:
            v = [fmt % x for x in r]
            f.write(f"({v[0]} ({v[1]} {v[2]} {v[3]}) ({v[4]} {v[5]} {v[6]}))\n")
        f.write(")\n")


def read_table(path):
    """Parse 6DoF.dat -> (N,7) array [t, tx,ty,tz, rx,ry,rz]."""
    with open(path, "rb") as f:
        buf = f.read()
    buf = re.sub(rb"//[^\n]*", b"", buf)
    m = re.match(rb"\s*(\d+)\s*\(", buf)
    if m is None:
        raise FoamError(f"{path}: expected 'N (' at the top of the motion table")
    n = int(m.group(1))
    a = np.array(buf[m.end() :].replace(b"(", b" ").replace(b")", b" ").split(), dtype=np.float64)
    if a.size != 7 * n:
        raise FoamError(f"{path}: table says {n} rows but holds {a.size} numbers (expected {7 * n})")
    rows = a.reshape(n, 7)
    if np.any(np.diff(rows[:, 0]) <= 0):
        raise FoamError(f"{path}: table times are not strictly increasing")
    return rows


def interpolate(rows, t):
    """Linear table lookup with clamping: Function1s::Table default bounds handling."""
    ts = rows[:, 0]
    if t <= ts[0]:
        return rows[0, 1:].copy()
    if t >= ts[-1]:
        return rows[-1, 1:].copy()
    i = int(np.searchsorted(ts, t, side="right")) - 1
    s = (t - ts[i]) / (ts[i + 1] - ts[i])
    return rows[i, 1:] + s * (rows[i + 1, 1:] - rows[i, 1:])


def rotation_matrix_xyz(deg):
    """quaternion(XYZ, angles): R = Rx * Ry * Rz, angles in degrees."""
    ax, ay, az = (math.radians(v) for v in deg)
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def transform_points(points0, rows, t, cofg=(0.0, 0.0, 0.0)):
    """p' = R (p - CofG) + CofG + translation."""
    v = interpolate(rows, t)
    R = rotation_matrix_xyz(v[3:6])
    c = np.asarray(cofg, float)
    return (points0 - c) @ R.T + c + v[0:3]
