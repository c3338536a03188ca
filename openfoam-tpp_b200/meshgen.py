"""polyMesh generators standing in for `gmsh` + `gmshToFoam` (absent from the build image).

The reference meshes its tanks with gmsh tets (circularSloshingTank/generate_mesh.py:17-51,
Makefile:60-62,73) and names the boundary `walls` / `atmosphere` and the cell zone
`internalMesh` (generate_mesh.py:29-51; constant/dynamicMeshDict:25).  These generators keep
that naming contract and produce meshes OpenFOAM itself would accept: owner < neighbour,
internal faces in upper-triangular order, boundary faces grouped by patch, face normals
pointing from owner to neighbour.

Cylinders are built as a ring-structured disc triangulation extruded in z into prisms,
optionally split into 3 tets per prism with the smallest-global-index diagonal rule
(conforming without Steiner points).  `cap` adds the spherical bottom of
generate_mesh.py:67-76 by mapping the column base onto the sphere z = -sqrt(R^2 - r^2).
The tutorial tank (sloshingTank3D6DoF) is a structured hex block whose y half-width follows
the chamfered profile described in SURVEY.md §8(d).
"""
from __future__ import annotations

import numpy as np

from .foamfile import PolyMesh


# ----------------------------------------------------------------------------------------
# generic cell-faces -> polyMesh
# ----------------------------------------------------------------------------------------
def build_polymesh(points, faces4, face_cell, patch_of_face, patch_names, patch_types=None, zone_name="internalMesh", extra_patch_info=None):
    """faces4: (N,4) int64 vertex ids per cell-face, oriented outward from its cell, -1 padded
    for triangles.  face_cell: (N,) owning cell of each copy.  patch_of_face(fc, fn, bf) ->
    patch index for boundary faces given centroids / unit normals / vertex ids (K,4).

    Returns a PolyMesh in OpenFOAM ordering."""
    faces4 = np.asarray(faces4, dtype=np.int64)
    face_cell = np.asarray(face_cell, dtype=np.int64)
    N = faces4.shape[0]
    nP = points.shape[0]
    srt = np.sort(np.where(faces4 < 0, np.iinfo(np.int64).max, faces4), axis=1)
    srt[srt == np.iinfo(np.int64).max] = nP  # sentinel beyond every point id
    base = nP + 1
    if base**3 < 2**62:
        # two int64 keys are always enough; one when the top vertex fits too
        k_lo = (srt[:, 0] * base + srt[:, 1]) * base + srt[:, 2]
        order = np.lexsort((srt[:, 3], k_lo))
        same = (k_lo[order][1:] == k_lo[order][:-1]) & (srt[order][1:, 3] == srt[order][:-1, 3])
    else:
        order = np.lexsort((srt[:, 3], srt[:, 2], srt[:, 1], srt[:, 0]))
        s = srt[order]
        same = np.all(s[1:] == s[:-1], axis=1)
    first = np.ones(N, dtype=bool)
    first[1:] = ~same  # first copy of each distinct face (in sorted order)
    paired = np.zeros(N, dtype=bool)
    paired[:-1] = same  # copy i (sorted) has its twin at i+1
    if np.any(same[1:] & same[:-1]):
        raise ValueError("a face is shared by more than two cells")
    i_first = order[first]
    is_int = paired[first]
    twin = order[np.nonzero(first)[0] + 1 - (~is_int)]  # for boundary faces twin == self
    c1 = face_cell[i_first]
    c2 = face_cell[twin]
    # internal
    ii = np.nonzero(is_int)[0]
    a, b = c1[ii], c2[ii]
    own = np.minimum(a, b)
    nei = np.maximum(a, b)
    src = np.where(a <= b, i_first[ii], twin[ii])  # owner's copy carries the orientation
    nC = int(face_cell.max()) + 1
    o = np.argsort(own * nC + nei, kind="stable")
    int_faces = faces4[src[o]]
    int_own, int_nei = own[o], nei[o]
    # boundary
    bi = i_first[~is_int]
    bf = faces4[bi]
    bc = face_cell[bi]
    fc, fn = _face_centroid_normal(points, bf)
    pid = np.asarray(patch_of_face(fc, fn, bf), dtype=np.int64)
    o = np.lexsort((bi, bc, pid))  # by patch, then owner cell, then generation order
    bf, bc, pid = bf[o], bc[o], pid[o]
    all_faces = np.concatenate([int_faces, bf], axis=0)
    owner = np.concatenate([int_own, bc])
    sizes = (all_faces >= 0).sum(axis=1)
    off = np.zeros(all_faces.shape[0] + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    labels = all_faces[all_faces >= 0]  # row-major keeps per-face order
    patches = []
    start = int_faces.shape[0]
    patch_types = patch_types or ["patch"] * len(patch_names)
    for k, nm in enumerate(patch_names):
        n = int((pid == k).sum())
        p = {"name": nm, "type": patch_types[k], "nFaces": n, "startFace": start}
        if extra_patch_info and nm in extra_patch_info:
            p.update(extra_patch_info[nm])
        patches.append(p)
        start += n
    zones = {zone_name: np.arange(nC, dtype=np.int32)} if zone_name else {}
    return PolyMesh(points, off, labels, owner, int_nei, patches, zones)


def _face_centroid_normal(points, faces4):
    """Cheap centroid / unit normal (vertex average, fan normal) for patch classification."""
    f = np.where(faces4 < 0, faces4[:, [0]], faces4)
    p = points[f]  # (N,4,3)
    cnt = (faces4 >= 0).sum(axis=1)[:, None]
    w = (faces4 >= 0)[:, :, None]
    c = (p * w).sum(axis=1) / cnt
    n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    tri = faces4[:, 3] < 0
    n2 = np.cross(p[:, 2] - p[:, 0], p[:, 3] - p[:, 0])
    n = np.where(tri[:, None], n, n + n2)
    n /= np.linalg.norm(n, axis=1)[:, None]
    return c, n


# ----------------------------------------------------------------------------------------
# disc triangulation
# ----------------------------------------------------------------------------------------
def disc_triangulation(R, n_rings):
    """Centre point + rings i=1..n_rings of 6i points at radius i*R/n_rings; consecutive rings
    are zipped by angle.  Returns (pts2d (P,2), tris (T,3) CCW).  T = 6 n_rings^2."""
    pts = [np.zeros((1, 2))]
    starts = [0, 1]
    for i in range(1, n_rings + 1):
        n = 6 * i
        th = 2.0 * np.pi * np.arange(n) / n
        r = R * i / n_rings
        pts.append(np.stack([r * np.cos(th), r * np.sin(th)], axis=1))
        starts.append(starts[-1] + n)
    pts = np.concatenate(pts, axis=0)
    tris = []
    # ring 0 (centre) to ring 1
    for k in range(6):
        tris.append((0, 1 + k, 1 + (k + 1) % 6))
    for i in range(1, n_rings):
        na, nb = 6 * i, 6 * (i + 1)
        sa, sb = starts[i], starts[i + 1]
        ia = ib = 0
        # exact rational angles: a_k = k/na, b_k = k/nb (in turns); advance the smaller next angle
        while ia < na or ib < nb:
            # next angles as fractions, compared with integers to stay deterministic
            adv_b = ib < nb and (ia >= na or (ib + 1) * na <= (ia + 1) * nb)
            if adv_b:
                tris.append((sa + ia % na, sb + ib % nb, sb + (ib + 1) % nb))
                ib += 1
            else:
                tris.append((sa + ia % na, sb + ib % nb, sa + (ia + 1) % na))
                ia += 1
    tris = np.array(tris, dtype=np.int64)
    # enforce CCW
    p = pts[tris]
    area2 = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    flip = area2 < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    return pts, tris


# orientation-preserving symmetries of a prism, indexed by the slot holding the smallest id
_PRISM_ROT = np.array(
    [
        [0, 1, 2, 3, 4, 5],
        [1, 2, 0, 4, 5, 3],
        [2, 0, 1, 5, 3, 4],
        [3, 5, 4, 0, 2, 1],
        [4, 3, 5, 1, 0, 2],
        [5, 4, 3, 2, 1, 0],
    ]
)


def prisms_to_tets(prisms):
    """(N,6) prisms [a b c | d e f] (d over a ...) -> (3N,4) tets, prism-major.  Each quad
    side is cut by the diagonal through its smallest global vertex id, so neighbouring prisms
    agree on the cut."""
    k = np.argmin(prisms, axis=1)
    v = np.take_along_axis(prisms, _PRISM_ROT[k], axis=1)
    a, b, c, d, e, f = (v[:, i] for i in range(6))
    opt1 = np.minimum(b, f) < np.minimum(c, e)
    t1 = np.where(opt1[:, None], np.stack([a, b, c, f], 1), np.stack([a, b, c, e], 1))
    t2 = np.where(opt1[:, None], np.stack([a, b, f, e], 1), np.stack([a, e, c, f], 1))
    t3 = np.stack([a, e, f, d], 1)
    return np.stack([t1, t2, t3], axis=1).reshape(-1, 4)


def _orient_tets(points, tets):
    p = points[tets]
    vol6 = np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0])
    neg = vol6 < 0
    tets = tets.copy()
    tets[neg] = tets[neg][:, [0, 2, 1, 3]]
    return tets


def _tet_faces(tets):
    n = tets.shape[0]
    p0, p1, p2, p3 = (tets[:, i] for i in range(4))
    m1 = np.full(n, -1, dtype=np.int64)
    f = np.stack(
        [
            np.stack([p0, p2, p1, m1], 1),
            np.stack([p0, p1, p3, m1], 1),
            np.stack([p1, p2, p3, m1], 1),
            np.stack([p0, p3, p2, m1], 1),
        ],
        axis=1,
    ).reshape(-1, 4)
    return f, np.repeat(np.arange(n), 4)


def _prism_faces(pr):
    n = pr.shape[0]
    a, b, c, d, e, f = (pr[:, i] for i in range(6))
    m1 = np.full(n, -1, dtype=np.int64)
    fs = np.stack(
        [
            np.stack([a, c, b, m1], 1),
            np.stack([d, e, f, m1], 1),
            np.stack([a, b, e, d], 1),
            np.stack([b, c, f, e], 1),
            np.stack([c, a, d, f], 1),
        ],
        axis=1,
    ).reshape(-1, 4)
    return fs, np.repeat(np.arange(n), 5)


def _hex_faces(hx):
    n = hx.shape[0]
    v = [hx[:, i] for i in range(8)]
    idx = [(0, 4, 7, 3), (1, 2, 6, 5), (0, 1, 5, 4), (3, 7, 6, 2), (0, 3, 2, 1), (4, 5, 6, 7)]
    fs = np.stack([np.stack([v[i] for i in q], 1) for q in idx], axis=1).reshape(-1, 4)
    return fs, np.repeat(np.arange(n), 6)


# ----------------------------------------------------------------------------------------
# tanks
# ----------------------------------------------------------------------------------------
def cylinder_mesh(H, D, n_rings, n_layers, geo="flat", cell="tet", k0=0, k1=None, proc=None):
    """Cylinder of height H, diameter D on z in [0,H] (generate_mesh.py:17-19), or the `cap`
    variant whose bottom is the sphere of radius D/2 centred at the origin
    (generate_mesh.py:67-76).  Patches `walls`, `atmosphere` (z = H); zone `internalMesh`.

    cell: 'tet' (3 per prism, like the reference's gmsh tets) or 'prism'.
    k0,k1: build only layers [k0,k1) of the extrusion — the slab one rank owns in the
    `simple` n=(1 1 N) z-decomposition; the cut planes become processor patches
    (`proc` = (myProcNo, lowerNeighbour or None, upperNeighbour or None)).
    """
    R = 0.5 * D
    k1 = n_layers if k1 is None else k1
    p2, tris = disc_triangulation(R, n_rings)
    n2 = p2.shape[0]
    nT = tris.shape[0]
    r2 = np.minimum((p2**2).sum(axis=1), R * R)
    zb = -np.sqrt(R * R - r2) if geo == "cap" else np.zeros(n2)
    if geo not in ("flat", "cap"):
        raise ValueError(f"unknown geometry '{geo}' (flat|cap)")
    nl = k1 - k0
    levels = np.arange(k0, k1 + 1) / n_layers  # s in [0,1]
    z = zb[None, :] * (1.0 - levels[:, None]) + H * levels[:, None]
    z[levels == 1.0] = H
    pts = np.empty((nl + 1, n2, 3))
    pts[:, :, 0] = p2[None, :, 0]
    pts[:, :, 1] = p2[None, :, 1]
    pts[:, :, 2] = z
    pts = pts.reshape(-1, 3)
    lay = np.arange(nl)[:, None, None] * n2
    pr = np.concatenate([tris[None] + lay, tris[None] + lay + n2], axis=2).reshape(-1, 6)
    # global vertex ids (for a slab-independent diagonal choice) = local + k0*n2
    if cell == "tet":
        tets = prisms_to_tets(pr + k0 * n2) - k0 * n2
        tets = _orient_tets(pts, tets)
        faces4, fcell = _tet_faces(tets)
    elif cell == "prism":
        faces4, fcell = _prism_faces(pr)
    else:
        raise ValueError(f"unknown cell type '{cell}' (tet|prism)")
    names, types, extra = ["walls", "atmosphere"], ["patch", "patch"], {}
    lo_id = hi_id = None
    if proc is not None and proc[1] is not None:
        names.append(f"procBoundary{proc[0]}to{proc[1]}")
        types.append("processor")
        extra[names[-1]] = {"myProcNo": proc[0], "neighbProcNo": proc[1]}
        lo_id = len(names) - 1
    if proc is not None and proc[2] is not None:
        names.append(f"procBoundary{proc[0]}to{proc[2]}")
        types.append("processor")
        extra[names[-1]] = {"myProcNo": proc[0], "neighbProcNo": proc[2]}
        hi_id = len(names) - 1
    lvl_of_pt = np.repeat(np.arange(nl + 1), n2)

    def classify(fc, fn, bf):
        # a boundary face whose vertices all sit on the slab's top (bottom) level is the
        # atmosphere or a processor cut; everything else is wall
        l = lvl_of_pt[np.where(bf < 0, bf[:, [0]], bf)]
        pid = np.zeros(bf.shape[0], dtype=np.int64)
        top = np.all(l == nl, axis=1)
        bot = np.all(l == 0, axis=1)
        if k1 == n_layers:
            pid[top] = 1
        elif hi_id is not None:
            pid[top] = hi_id
        if lo_id is not None:
            pid[bot] = lo_id
        return pid

    return build_polymesh(pts, faces4, fcell, classify, names, types, extra_patch_info=extra)


def box_mesh(nx, ny, nz, lo=(0, 0, 0), hi=(1, 1, 1), cell="hex", top_patch=None, wall_name="walls", ywidth=None):
    """Structured block.  cell='hex' or 'tet' (each hex -> 2 prisms -> 6 tets).
    top_patch: name of the z-max patch (None: every boundary face is `wall_name`).
    ywidth(z) -> half-width scaling of y about the block centre line (tutorial tank chamfers)."""
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    xs = np.linspace(lo[0], hi[0], nx + 1)
    ys = np.linspace(lo[1], hi[1], ny + 1)
    zs = np.linspace(lo[2], hi[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    if ywidth is not None:
        yc = 0.5 * (lo[1] + hi[1])
        half = 0.5 * (hi[1] - lo[1])
        Y = yc + (Y - yc) / half * ywidth(Z)
    pts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)

    def vid(i, j, k):
        return i + (nx + 1) * (j + (ny + 1) * k)

    K, J, I = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    hx = np.stack(
        [vid(I, J, K), vid(I + 1, J, K), vid(I + 1, J + 1, K), vid(I, J + 1, K), vid(I, J, K + 1), vid(I + 1, J, K + 1), vid(I + 1, J + 1, K + 1), vid(I, J + 1, K + 1)],
        axis=1,
    ).astype(np.int64)
    if cell == "hex":
        faces4, fcell = _hex_faces(hx)
    elif cell == "tet":
        # hex -> two prisms along the (0,2) diagonal of the bottom/top faces
        pa = hx[:, [0, 1, 2, 4, 5, 6]]
        pb = hx[:, [0, 2, 3, 4, 6, 7]]
        pr = np.stack([pa, pb], axis=1).reshape(-1, 6)
        tets = _orient_tets(pts, prisms_to_tets(pr))
        faces4, fcell = _tet_faces(tets)
    else:
        raise ValueError(cell)
    ztop = hi[2]
    eps = 1e-9 * np.abs(hi - lo).max()
    names = [wall_name] + ([top_patch] if top_patch else [])
    types = ["wall" if wall_name == "wall" else "patch"] + (["patch"] if top_patch else [])

    def classify(fc, fn, bf):
        pid = np.zeros(fc.shape[0], dtype=np.int64)
        if top_patch:
            pid[(np.abs(fc[:, 2] - ztop) < eps) & (fn[:, 2] > 0.999)] = 1
        return pid

    return build_polymesh(pts, faces4, fcell, classify, names, types, zone_name="internalMesh" if top_patch else "all")


def sloshing_tank3d_mesh(nx=10, ny=20, nz=15):
    """The closed tutorial tank of sloshingTank3D6DoF (single patch `wall`, zone `all`:
    sloshingTank3D6DoF/0/U:22, constant/dynamicMeshDict:25).  Its blockMeshDict is not in the
    reference (Allrun:7); geometry per SURVEY.md §8(d): depth 20 (x), width 40 (y), height 30,
    45-degree chamfers of height 5 (bottom) and 10 (top), shifted so the tank spans
    z in [-10, 20] and the free surface (setFieldsDict:26-29) sits at z = 0."""
    z0, z1 = -10.0, 20.0
    hb, ht, W = 5.0, 10.0, 20.0

    def half(z):
        w = np.full_like(z, W)
        lowc = z < z0 + hb
        w = np.where(lowc, W - (z0 + hb - z), w)
        upc = z > z1 - ht
        w = np.where(upc, W - (z - (z1 - ht)), w)
        return w

    # put grid planes exactly on the chamfer breaks when the resolution allows it
    return box_mesh(nx, ny, nz, lo=(-10.0, -W, z0), hi=(10.0, W, z1), cell="hex", top_patch=None, wall_name="wall", ywidth=half)


# ----------------------------------------------------------------------------------------
# geometry (numpy restatement used by setFields / tests; the solver and the oracle have
# their own)
# ----------------------------------------------------------------------------------------
def face_geometry(mesh: PolyMesh):
    """Face centres and area vectors, OpenFOAM primitiveMesh convention (triangle fan about
    the vertex average, area-weighted centroid)."""
    off, lab, P = mesh.face_offsets.astype(np.int64), mesh.face_labels, mesh.points
    nF = mesh.n_faces
    sizes = np.diff(off)
    Cf = np.zeros((nF, 3))
    Sf = np.zeros((nF, 3))
    for s in np.unique(sizes):
        idx = np.nonzero(sizes == s)[0]
        v = lab[off[idx][:, None] + np.arange(s)[None, :]]
        p = P[v]  # (n,s,3)
        if s == 3:
            Cf[idx] = (p[:, 0] + p[:, 1] + p[:, 2]) / 3.0
            Sf[idx] = 0.5 * np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
            continue
        fc = p.mean(axis=1)
        sumN = np.zeros((idx.size, 3))
        sumA = np.zeros(idx.size)
        sumAc = np.zeros((idx.size, 3))
        for i in range(s):
            a, b = p[:, i], p[:, (i + 1) % s]
            c = a + b + fc
            n = np.cross(b - a, fc - a)
            an = np.linalg.norm(n, axis=1)
            sumN += n
            sumA += an
            sumAc += an[:, None] * c
        Cf[idx] = sumAc / (3.0 * sumA[:, None])
        Sf[idx] = 0.5 * sumN
    return Cf, Sf


def cell_geometry(mesh: PolyMesh, Cf=None, Sf=None):
    """Cell centres / volumes, OpenFOAM primitiveMeshCellCentresAndVols convention."""
    if Cf is None:
        Cf, Sf = face_geometry(mesh)
    nC, nI = mesh.n_cells, mesh.n_internal
    own, nei = mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64)
    cEst = np.zeros((nC, 3))
    nf = np.zeros(nC)
    np.add.at(cEst, own, Cf)
    np.add.at(nf, own, 1)
    np.add.at(cEst, nei, Cf[:nI])
    np.add.at(nf, nei, 1)
    cEst /= nf[:, None]
    C = np.zeros((nC, 3))
    V = np.zeros(nC)
    pyr = np.einsum("ij,ij->i", Sf, Cf - cEst[own])
    pc = 0.75 * Cf + 0.25 * cEst[own]
    np.add.at(C, own, pyr[:, None] * pc)
    np.add.at(V, own, pyr)
    pyr = np.einsum("ij,ij->i", Sf[:nI], cEst[nei] - Cf[:nI])
    pc = 0.75 * Cf[:nI] + 0.25 * cEst[nei]
    np.add.at(C, nei, pyr[:, None] * pc)
    np.add.at(V, nei, pyr)
    C /= V[:, None]
    V /= 3.0
    return C, V


# ----------------------------------------------------------------------------------------
# unstructured tets (stand-in for `gmsh -3 cylinder.geo`, generate_mesh.py:17-51: Delaunay tets of
# uniform size lc, in no particular cell order)
# ----------------------------------------------------------------------------------------
def _tet_quality(p, t):
    """(volume, radius-ratio-like quality in [0,1]: 1 for the regular tet)."""
    a = p[t]
    v = np.einsum("ij,ij->i", np.cross(a[:, 1] - a[:, 0], a[:, 2] - a[:, 0]), a[:, 3] - a[:, 0]) / 6.0
    e2 = np.zeros(t.shape[0])
    for i in range(4):
        for j in range(i + 1, 4):
            e2 += ((a[:, i] - a[:, j]) ** 2).sum(axis=1)
    return v, 6.0 * np.sqrt(2.0) * np.abs(v) / np.maximum((e2 / 6.0) ** 1.5, 1e-300)


def unstructured_cylinder_tets(H, D, lc, seed=0, iters=60, q_min=0.12):
    """Points and tetrahedra of an unstructured Delaunay mesh of the flat-bottom cylinder
    (0 <= z <= H, r <= D/2) with target edge length lc: boundary nodes fixed on rings, interior
    nodes relaxed with Persson-Strang spring forces between Delaunay retriangulations, then
    slivers removed by local perturbation.  Needs scipy."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    R = 0.5 * D
    h = lc * 1.21  # node spacing for which the cell count matches gmsh's at this lc (41 895 tets at 9 mm, G1)
    # fixed boundary nodes: wall rings, and concentric rings on both end discs
    nz = max(int(round(H / h)), 1)
    nth = max(int(round(2 * np.pi * R / h)), 8)
    bnd = []
    for k in range(nz + 1):
        th = 2 * np.pi * (np.arange(nth) + 0.5 * (k % 2)) / nth
        bnd.append(np.c_[R * np.cos(th), R * np.sin(th), np.full(nth, H * k / nz)])
    nr = max(int(round(R / h)), 1)
    for z in (0.0, H):
        for i in range(nr):
            r = R * i / nr
            n = max(int(round(2 * np.pi * r / h)), 1)
            th = 2 * np.pi * (np.arange(n) + 0.37 * i) / n
            bnd.append(np.c_[r * np.cos(th), r * np.sin(th), np.full(n, z)])
    bnd = np.concatenate(bnd)
    nb = bnd.shape[0]
    # interior nodes: jittered body-centred lattice clipped half a spacing inside
    a = h * 2.0 / np.sqrt(3.0) * 0.97
    g = np.arange(-R, R + a, a)
    gz = np.arange(0.0, H + a, a)
    X, Y, Z = np.meshgrid(g, g, gz, indexing="ij")
    lat = np.c_[X.ravel(), Y.ravel(), Z.ravel()]
    lat = np.concatenate([lat, lat + 0.5 * a])
    lat += rng.uniform(-0.3 * a, 0.3 * a, lat.shape)  # destroys the lattice: the relaxed mesh is irregular
    m = (np.hypot(lat[:, 0], lat[:, 1]) < R - 0.55 * h) & (lat[:, 2] > 0.55 * h) & (lat[:, 2] < H - 0.55 * h)
    p = np.concatenate([bnd, lat[m]])

    def inside(q, s):
        r = np.hypot(q[:, 0], q[:, 1])
        sc = np.minimum((R - s) / np.maximum(r, 1e-300), 1.0)
        q[:, 0] *= sc
        q[:, 1] *= sc
        q[:, 2] = np.clip(q[:, 2], s, H - s)

    def triangulate(p):
        t = Delaunay(p).simplices.astype(np.int64)
        v, q = _tet_quality(p, t)
        # hull slivers: four boundary nodes, (almost) no volume
        flat = (t < nb).all(axis=1) & (q < 0.05)
        return t[~flat], q[~flat]

    t, q = triangulate(p)
    for it in range(iters):
        e = np.concatenate([t[:, [i, j]] for i in range(4) for j in range(i + 1, 4)])
        e = np.unique(np.sort(e, axis=1), axis=0)
        d = p[e[:, 0]] - p[e[:, 1]]
        L = np.linalg.norm(d, axis=1)
        L0 = 1.15 * np.sqrt((L**2).mean())
        F = np.maximum(L0 - L, 0.0)
        fv = (F / np.maximum(L, 1e-300))[:, None] * d
        tot = np.zeros_like(p)
        np.add.at(tot, e[:, 0], fv)
        np.add.at(tot, e[:, 1], -fv)
        tot[:nb] = 0.0
        p = p + 0.15 * tot
        inside(p[nb:], 0.35 * h)
        if it % 4 == 3 or it == iters - 1:
            t, q = triangulate(p)
    # slivers: jiggle one interior node of each bad tet and retriangulate
    for _ in range(40):
        bad = np.nonzero(q < q_min)[0]
        if bad.size == 0:
            break
        nodes = np.unique(t[bad].ravel())
        nodes = nodes[nodes >= nb]
        if nodes.size == 0:
            break
        p[nodes] += rng.normal(0.0, 0.12 * h, (nodes.size, 3))
        inside(p[nb:], 0.35 * h)
        t, q = triangulate(p)
    used = np.unique(t.ravel())
    lut = np.full(p.shape[0], -1, dtype=np.int64)
    lut[used] = np.arange(used.size)
    return p[used], lut[t]


def unstructured_cylinder_mesh(H, D, lc, seed=0, iters=60):
    """polyMesh of `unstructured_cylinder_tets` with the reference's naming (patches `walls`,
    `atmosphere`, zone `internalMesh`); cells in the triangulator's order, like a gmshToFoam mesh."""
    pts, tets = unstructured_cylinder_tets(H, D, lc, seed, iters)
    tets = _orient_tets(pts, tets)
    faces4, fcell = _tet_faces(tets)
    eps = 1e-9 * max(H, D)

    def classify(fc, fn, bf):
        z = pts[bf[:, :3]][:, :, 2]
        return (np.abs(z - H).max(axis=1) < eps).astype(np.int64)

    return build_polymesh(pts, faces4, fcell, classify, ["walls", "atmosphere"], ["patch", "patch"])


def renumbered(mesh: PolyMesh, new_of_old):
    """The same mesh with cell `c` renamed `new_of_old[c]`, in valid OpenFOAM order again: faces
    whose owner would exceed their neighbour are flipped (vertex order reversed), internal faces
    re-sorted upper-triangular, boundary faces keep their patch and their order inside it."""
    new_of_old = np.asarray(new_of_old, dtype=np.int64)
    nI, nF = mesh.n_internal, mesh.n_faces
    off, lab = mesh.face_offsets.astype(np.int64), mesh.face_labels.astype(np.int64)
    o = new_of_old[mesh.owner.astype(np.int64)]
    n = new_of_old[mesh.neighbour.astype(np.int64)]
    flip = np.zeros(nF, dtype=bool)
    flip[:nI] = o[:nI] > n
    own_i, nei_i = np.minimum(o[:nI], n), np.maximum(o[:nI], n)
    order = np.concatenate([np.lexsort((nei_i, own_i)), np.arange(nI, nF)])
    cnt = np.diff(off)
    # vertex lists in the new face order; a flipped face keeps its first vertex and reverses the rest
    starts = off[:-1][order]
    c2 = cnt[order]
    new_off = np.zeros(nF + 1, dtype=np.int64)
    new_off[1:] = np.cumsum(c2)
    pos = np.arange(new_off[-1]) - np.repeat(new_off[:-1], c2)
    fl = np.repeat(flip[order], c2)
    cc = np.repeat(c2, c2)
    src = np.repeat(starts, c2) + np.where(fl & (pos > 0), cc - pos, pos)
    new_lab = lab[src]
    owner = np.concatenate([own_i, o[nI:]])[order]
    neighbour = nei_i[order[:nI]]
    zones = {k: np.sort(new_of_old[np.asarray(v, dtype=np.int64)]).astype(np.int32) for k, v in (mesh.cell_zones or {}).items()}
    return PolyMesh(mesh.points, new_off, new_lab, owner, neighbour, [dict(p) for p in mesh.patches], zones)


def shuffled(mesh: PolyMesh, seed=0):
    """Random cell order (what gmsh + gmshToFoam leave: no spatial coherence between neighbouring
    cell labels) - the worst case for the gathers; the reference never calls renumberMesh
    (circularSloshingTank/Makefile:71-86)."""
    return renumbered(mesh, np.random.default_rng(seed).permutation(mesh.n_cells))
