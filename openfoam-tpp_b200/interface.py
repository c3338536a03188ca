"""VTK-free interface diagnostics (the parity metric of SURVEY.md §4 / §8f-3).

The reference extracts the alpha = 0.5 iso-surface with PyVista (cell data averaged to the points,
then a contour filter) and reduces it to `interface_summary.csv` (time,max_z,min_z,mean_z,
num_points; main.py:751-780) and to a wall elevation series (main.py:784-798).  PyVista/VTK are
not available here; `cell_to_point` + `iso_points` restate those two filters on the mesh edges
(any cell type), `extract_interface` writes the same two CSV files, `iso_wall_mode1` / `beat_fit`
reduce the iso-surface to the first azimuthal sloshing mode the way the golden series G4 was
reduced.  `ColumnSampler` is an independent second measure for extruded meshes only: the water
column height h(x, y) = (1/A_col) sum_col alpha V in the tank frame.
"""
from __future__ import annotations

import numpy as np

from . import meshgen


class ColumnSampler:
    """Columns of an extruded mesh (cylinder_mesh): cells sharing a base triangle."""

    def __init__(self, mesh, z_bottom=0.0):
        H = float(mesh.points[:, 2].max() - mesh.points[:, 2].min())
        C, V = meshgen.cell_geometry(mesh)
        self.V = V
        # column id from the horizontal position of the cell centre: cells of one prism column
        # (3 tets per prism, all layers) share the base triangle; identify columns by rounding
        # the centroid of the *prism* = mean over its tets is not available, so group by the
        # triangle that contains the cell centre -> use the generator's numbering when present
        n_layers, n_tri, per = mesh_layout(mesh, C)
        self.n_layers, self.n_tri, self.per = n_layers, n_tri, per
        ids = np.arange(mesh.n_cells)
        self.col = (ids % (n_tri * per)) // per
        area = np.zeros(n_tri)
        np.add.at(area, self.col, V)
        self.H = H
        self.area = area / H  # column volume / tank height = base area
        xy = np.zeros((n_tri, 2))
        np.add.at(xy, self.col, C[:, :2] * V[:, None])
        vol = np.zeros(n_tri)
        np.add.at(vol, self.col, V)
        self.xy = xy / vol[:, None]
        self.r = np.hypot(self.xy[:, 0], self.xy[:, 1])
        self.theta = np.arctan2(self.xy[:, 1], self.xy[:, 0])
        self.z_bottom = z_bottom

    def heights(self, alpha):
        w = np.zeros(self.n_tri)
        np.add.at(w, self.col, alpha * self.V)
        return self.z_bottom + w / self.area

    def summary(self, alpha):
        h = self.heights(alpha)
        return float(h.max()), float(h.min()), float((h * self.area).sum() / self.area.sum())

    def wall_mode1(self, alpha, r_frac=0.85):
        """Least-squares fit z = z0 + C cos(theta) + S sin(theta) on the columns near the wall:
        amplitude and phase of the first azimuthal sloshing mode (tank frame)."""
        h = self.heights(alpha)
        m = self.r > r_frac * self.r.max()
        A = np.stack([np.ones(m.sum()), np.cos(self.theta[m]), np.sin(self.theta[m])], axis=1)
        wgt = np.sqrt(self.area[m])
        z0, c, s = np.linalg.lstsq(A * wgt[:, None], h[m] * wgt, rcond=None)[0]
        return float(np.hypot(c, s)), float(np.arctan2(s, c)), float(z0)


def mesh_layout(mesh, C=None):
    """(n_layers, n_triangles, cells per prism) of a cylinder_mesh: recovered from the point
    count (points = (n_layers+1) * n2) and the cell count."""
    z = np.unique(np.round(mesh.points[:, 2], 12))
    # flat tanks: every level is a z-plane
    n_levels = z.size
    n_layers = n_levels - 1
    n2 = mesh.n_points // n_levels
    if n2 * n_levels != mesh.n_points:
        raise ValueError("ColumnSampler needs a flat-bottom extruded mesh (cylinder_mesh geo='flat')")
    per_layer = mesh.n_cells // n_layers
    # n_tri from Euler: rings structure gives T = 6 nr^2, P2 = 1 + 3 nr (nr + 1)
    nr = int(round((-3 + np.sqrt(9 + 12 * (n2 - 1))) / 6))
    n_tri = 6 * nr * nr
    per = per_layer // n_tri
    if per * n_tri * n_layers != mesh.n_cells:
        raise ValueError("unexpected cell layout")
    return n_layers, n_tri, per


# ---- the reference's own metric, without VTK -------------------------------------------------------
def cell_to_point(mesh, cell_values):
    """vtkCellDataToPointData: a point takes the plain average of the cells that use it."""
    off, lab = mesh.face_offsets.astype(np.int64), mesh.face_labels.astype(np.int64)
    cnt = np.diff(off)
    nI = mesh.n_internal
    fc_own = np.repeat(mesh.owner.astype(np.int64), cnt)
    pairs = [fc_own * mesh.n_points + lab]
    if nI:
        sel = np.repeat(np.arange(mesh.n_faces) < nI, cnt)
        fc_nei = np.repeat(np.concatenate([mesh.neighbour.astype(np.int64), np.zeros(mesh.n_faces - nI, dtype=np.int64)]), cnt)
        pairs.append(fc_nei[sel] * mesh.n_points + lab[sel])
    key = np.unique(np.concatenate(pairs))  # distinct (cell, point) incidences
    cell, point = key // mesh.n_points, key % mesh.n_points
    s = np.zeros(mesh.n_points)
    n = np.zeros(mesh.n_points)
    np.add.at(s, point, np.asarray(cell_values, dtype=np.float64)[cell])
    np.add.at(n, point, 1.0)
    return s / np.maximum(n, 1.0)


def mesh_edges(mesh):
    """distinct point pairs joined by a face edge (= every cell edge of tets, prisms and hexes)"""
    off, lab = mesh.face_offsets.astype(np.int64), mesh.face_labels.astype(np.int64)
    cnt = np.diff(off)
    pos = np.arange(lab.size) - np.repeat(off[:-1], cnt)
    nxt = np.where(pos + 1 < np.repeat(cnt, cnt), np.arange(lab.size) + 1, np.repeat(off[:-1], cnt))
    a, b = lab, lab[nxt]
    key = np.unique(np.minimum(a, b) * mesh.n_points + np.maximum(a, b))
    return key // mesh.n_points, key % mesh.n_points


def iso_points(mesh, points, point_values, iso=0.5, edges=None):
    """The points of the iso-surface a contour filter produces on linear cells: one per mesh edge
    whose end values straddle `iso`, linearly interpolated along the edge."""
    a, b = edges if edges is not None else mesh_edges(mesh)
    va, vb = point_values[a], point_values[b]
    cut = (va >= iso) != (vb >= iso)
    a, b, va, vb = a[cut], b[cut], va[cut], vb[cut]
    t = (iso - va) / (vb - va)
    return points[a] + t[:, None] * (points[b] - points[a])


def iso_wall_mode1(pts, centre_xy=(0.0, 0.0), R=0.1, r_frac=0.9):
    """First azimuthal mode of the interface at the wall, from iso-surface points: least-squares
    fit z = z0 + C cos(theta) + S sin(theta) over the points with tank-frame radius > r_frac R
    (theta about the tank centre `centre_xy`).  This is the fit SURVEY.md §4 applies to the
    reference's committed iso-surfaces (golden G4, tests/golden/g4_m1_series.csv), so a run and the
    golden are reduced by the same arithmetic.  -> (amplitude, phase, z0, n_points)"""
    x, y = pts[:, 0] - centre_xy[0], pts[:, 1] - centre_xy[1]
    m = np.hypot(x, y) > r_frac * R
    if m.sum() < 3:
        return 0.0, 0.0, 0.0, int(m.sum())
    th = np.arctan2(y[m], x[m])
    A = np.stack([np.ones(m.sum()), np.cos(th), np.sin(th)], axis=1)
    z0, c, s = np.linalg.lstsq(A, pts[m, 2], rcond=None)[0]
    return float(np.hypot(c, s)), float(np.arctan2(s, c)), float(z0), int(m.sum())


def beat_fit(t, amp, phase, f_forcing, t0=2.0, t1=None):
    """Reduces an m = 1 series (amplitude, phase in the tank frame, as `iso_wall_mode1` gives them)
    to the four numbers that describe a forced, lightly damped sloshing mode after the ramp:
    q(t) = a_f exp(i w t) + a_n exp((i w0 - gamma) t).  Linear least squares in (a_f, a_n) inside a
    scan over (w0, gamma).  -> dict(f0 [Hz], gamma [1/s], A_forced [m], A_free [m] at t0, rms [m])"""
    t = np.asarray(t, dtype=float)
    q = np.asarray(amp) * np.exp(1j * np.asarray(phase))
    m = t >= t0
    if t1 is not None:
        m &= t <= t1
    t, q = t[m], q[m]
    w = 2 * np.pi * f_forcing

    def solve(f0, g):
        B = np.stack([np.exp(1j * w * t), np.exp((1j * 2 * np.pi * f0 - g) * (t - t0))], axis=1)
        c, *_ = np.linalg.lstsq(B, q, rcond=None)
        return c, float(np.sqrt(np.mean(np.abs(B @ c - q) ** 2)))

    best = None
    f_grid, g_grid = np.linspace(f_forcing + 0.05, f_forcing + 0.6, 111), np.linspace(0.0, 0.6, 31)
    for _ in range(3):
        for f0 in f_grid:
            for g in g_grid:
                c, r = solve(f0, g)
                if best is None or r < best[0]:
                    best = (r, f0, g, c)
        df, dg = f_grid[1] - f_grid[0], g_grid[1] - g_grid[0]
        f_grid = np.linspace(best[1] - df, best[1] + df, 21)
        g_grid = np.linspace(max(best[2] - dg, 0.0), best[2] + dg, 21)
    r, f0, g, c = best
    return {"f0": float(f0), "gamma": float(g), "A_forced": float(abs(c[0])), "A_free": float(abs(c[1])), "rms": r}


def tet_points(mesh):
    """(nC, 4) point labels of every cell of a tetrahedral mesh, or None when some cell is not a tet."""
    off, lab = mesh.face_offsets.astype(np.int64), mesh.face_labels.astype(np.int64)
    cnt = np.diff(off)
    if np.any(cnt != 3):
        return None
    nI = mesh.n_internal
    cells = np.concatenate([np.repeat(mesh.owner.astype(np.int64), 3), np.repeat(mesh.neighbour.astype(np.int64), 3)])
    pts = np.concatenate([lab, lab[: 3 * nI]])
    key = np.unique(cells * mesh.n_points + pts)
    if key.size != 4 * mesh.n_cells or np.any(np.bincount(key // mesh.n_points, minlength=mesh.n_cells) != 4):
        return None
    return (key % mesh.n_points).reshape(mesh.n_cells, 4)


def iso_surface(mesh, points, point_values, iso=0.5, edges=None, tets=None):
    """Contour points AND triangles (marching tetrahedra) of a tetrahedral mesh: what the contour
    filter of the reference produces on its gmsh tets (main.py:770).  -> (pts (N,3), tris (M,3) into
    pts, (edge end a, edge end b, weight t) of every point for interpolating other fields).  On a
    non-tetrahedral mesh tris is empty."""
    a, b = edges if edges is not None else mesh_edges(mesh)
    va, vb = point_values[a], point_values[b]
    cut = (va >= iso) != (vb >= iso)
    ca, cb = a[cut], b[cut]
    t = (iso - va[cut]) / (vb[cut] - va[cut])
    pts = points[ca] + t[:, None] * (points[cb] - points[ca])
    tets = tet_points(mesh) if tets is None else tets
    if tets is None or pts.shape[0] == 0:
        return pts, np.zeros((0, 3), dtype=np.int64), (ca, cb, t)
    nP = mesh.n_points
    ekey = ca * nP + cb  # ascending (edges come sorted from mesh_edges, a < b)

    def eid(p, q):
        lo, hi = np.minimum(p, q), np.maximum(p, q)
        return np.searchsorted(ekey, lo * nP + hi)

    above = point_values[tets] >= iso
    k = above.sum(axis=1)
    tris = []
    for lone_above in (True, False):  # one vertex on its own side: a triangle
        sel = np.nonzero(k == (1 if lone_above else 3))[0]
        if sel.size:
            tv = tets[sel]
            m = above[sel] == lone_above
            lone = tv[m]
            rest = tv[~m].reshape(-1, 3)
            tris.append(np.stack([eid(lone, rest[:, 0]), eid(lone, rest[:, 1]), eid(lone, rest[:, 2])], axis=1))
    sel = np.nonzero(k == 2)[0]  # two and two: a quadrilateral, split into two triangles
    if sel.size:
        tv = tets[sel]
        m = above[sel]
        up = tv[m].reshape(-1, 2)
        dn = tv[~m].reshape(-1, 2)
        q0, q1, q2, q3 = eid(up[:, 0], dn[:, 0]), eid(up[:, 0], dn[:, 1]), eid(up[:, 1], dn[:, 1]), eid(up[:, 1], dn[:, 0])
        tris.append(np.stack([q0, q1, q2], axis=1))
        tris.append(np.stack([q0, q2, q3], axis=1))
    tris = np.concatenate(tris) if tris else np.zeros((0, 3), dtype=np.int64)
    return pts, tris, (ca, cb, t)


def write_vtp(path, pts, tris, point_data=None):
    """A VTK XML PolyData file (ASCII arrays) of the contour: the `interface_t<time>.vtp` files
    extract_interface leaves next to its CSVs (main.py:772-774), readable by ParaView / PyVista."""
    point_data = point_data or {}

    def arr(a, typ, name=None, nc=None):
        a = np.asarray(a)
        head = f'<DataArray type="{typ}"' + (f' Name="{name}"' if name else "") + (f' NumberOfComponents="{nc}"' if nc else "") + ' format="ascii">'
        fmt = "%d" if typ.startswith("Int") else "%.9g"
        return head + "\n" + " ".join(fmt % v for v in a.reshape(-1)) + "\n</DataArray>\n"

    n, m = int(pts.shape[0]), int(tris.shape[0])
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="PolyData" version="0.1" byte_order="LittleEndian">\n<PolyData>\n')
        f.write(f'<Piece NumberOfPoints="{n}" NumberOfVerts="{0 if m else n}" NumberOfLines="0" NumberOfStrips="0" NumberOfPolys="{m}">\n')
        scal = [k for k, v in point_data.items() if np.asarray(v).ndim == 1]
        vec = [k for k, v in point_data.items() if np.asarray(v).ndim == 2]
        f.write("<PointData" + (f' Scalars="{scal[0]}"' if scal else "") + (f' Vectors="{vec[0]}"' if vec else "") + ">\n")
        for k, v in point_data.items():
            v = np.asarray(v, dtype=np.float32)
            f.write(arr(v, "Float32", k, v.shape[1] if v.ndim == 2 else None))
        f.write("</PointData>\n<Points>\n" + arr(np.asarray(pts, dtype=np.float32), "Float32", "Points", 3) + "</Points>\n")
        if m:
            f.write("<Polys>\n" + arr(tris, "Int64", "connectivity") + arr(3 * np.arange(1, m + 1), "Int64", "offsets") + "</Polys>\n")
        else:
            f.write("<Verts>\n" + arr(np.arange(n), "Int64", "connectivity") + arr(np.arange(1, n + 1), "Int64", "offsets") + "</Verts>\n")
        f.write("</Piece>\n</PolyData>\n</VTKFile>\n")


def extract_interface(case_dir, r_target=None, write=True, vtp=False):
    """VTK-free restatement of the reference's `extract_interface` (main.py:727-818): for every
    time directory the alpha.water = 0.5 iso-surface of the point-averaged field on the mesh at
    that time (`<time>/polyMesh/points`: lab frame), reduced to
    postProcessing/interface/interface_summary.csv (time,max_z,min_z,mean_z,num_points) and
    wall_elevation.csv (time,theta,zeta_wall; points with r > 0.98 R in 64 theta bins).
    vtp=True also writes the per-frame `interface_t<time>.vtp` files (main.py:772-774): contour points,
    triangles (tetrahedral meshes) and the point data alpha.water / p_rgh / U / p / rho interpolated along
    the cut edges.  Returns the summary rows."""
    import os
    import re

    from . import foamfile as ff

    mesh = ff.read_polymesh(case_dir)
    edges = mesh_edges(mesh)
    tets = tet_points(mesh) if vtp else None
    if r_target is None:
        m = re.search(r"_D([\d.]+)_", os.path.basename(os.path.normpath(case_dir)))
        r_target = float(m.group(1)) / 2.0 if m else 0.1
    summary, wall = ["time,max_z,min_z,mean_z,num_points"], ["time,theta,zeta_wall"]
    rows = []
    bins = np.linspace(-np.pi, np.pi, 65)
    for t, name in ff.time_dirs(case_dir):
        fp = os.path.join(case_dir, name, "alpha.water")
        if not os.path.exists(fp):
            continue
        alpha = ff.read_field(fp).internal_array(mesh.n_cells)
        pp = os.path.join(case_dir, name, "polyMesh", "points")
        pts_mesh = ff.read_points(pp) if os.path.exists(pp) else mesh.points
        pa = cell_to_point(mesh, alpha)
        pts = iso_points(mesh, pts_mesh, pa, 0.5, edges)
        if vtp and write:
            p2, tris, (ca, cb, tt) = iso_surface(mesh, pts_mesh, pa, 0.5, edges, tets)
            data = {"alpha.water": np.full(len(p2), 0.5)}
            for fld in ("p_rgh", "p", "rho", "U"):
                fq = os.path.join(case_dir, name, fld)
                if os.path.exists(fq):
                    pv = ff.read_field(fq).internal_array(mesh.n_cells)
                    pv = np.stack([cell_to_point(mesh, pv[:, k]) for k in range(3)], axis=1) if pv.ndim == 2 else cell_to_point(mesh, pv)
                    data[fld] = pv[ca] + (tt[:, None] if pv.ndim == 2 else tt) * (pv[cb] - pv[ca])
            out = os.path.join(case_dir, "postProcessing", "interface")
            os.makedirs(out, exist_ok=True)
            write_vtp(os.path.join(out, f"interface_t{t:.6f}.vtp"), p2, tris, data)
        if len(pts) == 0:
            summary.append(f"{t},0,0,0,0")
            rows.append((t, 0.0, 0.0, 0.0, 0))
            continue
        z = pts[:, 2]
        summary.append(f"{t},{z.max()},{z.min()},{z.mean()},{len(pts)}")
        rows.append((t, float(z.max()), float(z.min()), float(z.mean()), len(pts)))
        r = np.hypot(pts[:, 0], pts[:, 1])
        wm = r > r_target * 0.98
        if wm.any():
            wp = pts[wm]
            th = np.arctan2(wp[:, 1], wp[:, 0])
            which = np.digitize(th, bins) - 1
            for b in range(64):
                sel = which == b
                if sel.any():
                    wall.append(f"{t},{(bins[b] + bins[b + 1]) / 2.0},{wp[sel, 2].mean()}")
    if write:
        out = os.path.join(case_dir, "postProcessing", "interface")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "interface_summary.csv"), "w") as f:
            f.write("\n".join(summary))
        with open(os.path.join(out, "wall_elevation.csv"), "w") as f:
            f.write("\n".join(wall))
    return rows
