"""VTK-free interface diagnostics (the parity metric of SURVEY.md §4 / §8f-3).

The reference extracts the alpha = 0.5 iso-surface with PyVista and reduces it to
`interface_summary.csv` (time,max_z,min_z,mean_z,num_points; main.py:751-780) and to a wall
elevation series (main.py:784-798).  PyVista/VTK are not available here, and the quantity that
matters for parity is the free-surface elevation, so this module measures it directly from the
cell data: the water column height h(x, y) = (1/A_col) * sum_col alpha V over each vertical
column of the extruded tank mesh, in the tank frame.  For a single-valued interface this is the
iso-surface elevation up to the O(cell) smearing of the VOF front.
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
