"""VTK-free interface extraction (reference main.py:727-818): the points a
contour filter puts on the alpha = 0.5 surface are the straddling mesh edges, interpolated."""
import os

import numpy as np

from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import foamfile as ff
from openfoam_tpp_b200 import interface as itf
from openfoam_tpp_b200 import meshgen as mg


def test_iso_points_of_a_linear_field_lie_on_its_plane():
    mesh = mg.cylinder_mesh(0.05, 0.04, 4, 6, "flat", "tet")
    n = np.array([0.3, -0.2, 1.0])
    d0 = 0.021
    f = lambda x: 0.5 - 40.0 * (x @ n - d0)  # 0.5 exactly on the plane n.x = d0
    pv = f(mesh.points)
    pts = itf.iso_points(mesh, mesh.points, pv)
    assert len(pts) > 50
    assert np.abs(pts @ n - d0).max() < 1e-14
    # one point per cut edge
    a, b = itf.mesh_edges(mesh)
    assert len(pts) == int(((pv[a] >= 0.5) != (pv[b] >= 0.5)).sum())
    # tets: 6 edges per cell, Euler-consistent edge count (V - E + F - C = 1 for a ball)
    assert mesh.n_points - len(a) + mesh.n_faces - mesh.n_cells == 1


def test_cell_to_point_is_the_plain_average():
    mesh = mg.box_mesh(2, 2, 2)
    v = np.arange(mesh.n_cells, dtype=float)
    pv = itf.cell_to_point(mesh, v)
    centre = np.argmin(np.abs(mesh.points - 0.5).sum(axis=1))  # shared by all 8 hexes
    assert pv[centre] == v.mean()
    corner = np.argmin(np.abs(mesh.points).sum(axis=1))        # belongs to one hex
    C, _ = mg.cell_geometry(mesh)
    assert pv[corner] == v[np.argmin(np.abs(C - 0.25).sum(axis=1))]


def test_extract_interface_files(tmp_path):
    d = str(tmp_path / "case_H0.004_D0.0221_flat_R0.005_f2.0")
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=6, n_layers=8)
    rows = itf.extract_interface(d)
    assert len(rows) == 1 and rows[0][0] == 0.0
    t, zmax, zmin, zmean, npts = rows[0]
    # flat fill at H/2 (setFields box): the smeared point field crosses 0.5 within one layer of it
    assert npts > 100 and abs(zmean - 0.002) < 0.004 / 8 and zmax - zmin <= 2 * 0.004 / 8
    out = os.path.join(d, "postProcessing", "interface")
    head = open(os.path.join(out, "interface_summary.csv")).read().splitlines()
    assert head[0] == "time,max_z,min_z,mean_z,num_points" and len(head) == 2
    wall = open(os.path.join(out, "wall_elevation.csv")).read().splitlines()
    assert wall[0] == "time,theta,zeta_wall" and 1 < len(wall) <= 65


def test_marching_tets_surface_and_vtp(tmp_path):
    """The contour as a triangulated surface (what PyVista's contour filter gives the reference on its
    tets, main.py:770-774): for a planar field the triangles tile the plane's cut through the
    cylinder (area of the ellipse to the accuracy of the polygonal rim), every triangle lies in the
    plane, and the .vtp file carries the points, the triangles and the point data."""
    import xml.etree.ElementTree as ET

    from openfoam_tpp_b200 import interface as it
    from openfoam_tpp_b200 import meshgen as mg

    R, H = 0.1, 0.208
    mesh = mg.cylinder_mesh(H, 2 * R, 8, 12, "flat", "tet")
    n = np.array([0.2, -0.1, 1.0])
    pv = 0.5 + (0.104 - mesh.points @ n)  # a plane through z = 0.104 at the axis, tilted
    pts, tris, (ca, cb, t) = it.iso_surface(mesh, mesh.points, pv, 0.5)
    assert len(pts) == len(it.iso_points(mesh, mesh.points, pv, 0.5)) and len(tris) > len(pts)
    assert np.abs(pts @ n - 0.104).max() < 1e-12
    a, b, c = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1).sum()
    exact = np.pi * R * R * np.linalg.norm(n) / n[2]
    assert abs(area / exact - 1) < 0.03, (area, exact)
    f = tmp_path / "interface_t0.000000.vtp"
    it.write_vtp(str(f), pts, tris, {"alpha.water": np.full(len(pts), 0.5), "U": np.zeros((len(pts), 3))})
    piece = ET.parse(str(f)).getroot().find("PolyData/Piece")
    assert int(piece.get("NumberOfPoints")) == len(pts) and int(piece.get("NumberOfPolys")) == len(tris)
    names = [d.get("Name") for d in piece.find("PointData")]
    assert names == ["alpha.water", "U"]
    conn = np.array(piece.find("Polys/DataArray[@Name='connectivity']").text.split(), dtype=np.int64)
    assert conn.size == 3 * len(tris) and conn.max() < len(pts)
    # prisms: points only (vertices)
    pm = mg.cylinder_mesh(H, 2 * R, 4, 4, "flat", "prism")
    p2, t2, _ = it.iso_surface(pm, pm.points, 0.5 + (0.104 - pm.points @ n), 0.5)
    assert len(p2) > 0 and len(t2) == 0
