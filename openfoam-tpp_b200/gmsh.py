"""Gmsh msh 2.2 (ASCII) -> polyMesh: the `gmshToFoam cylinder.msh` step of the reference's run
recipe (circularSloshingTank/Makefile:73) for the meshes generate_mesh.py describes: tetrahedra
(element type 4) in one physical volume (`internalMesh` -> cellZone), boundary triangles
(type 2) tagged with physical surfaces (`atmosphere`, `walls` -> patches of type patch)
(generate_mesh.py:29-51, `Mesh.MshFileVersion = 2.2`, main.py:304-308 `-format msh2`).

Cells keep the order of the $Elements section; faces are produced in OpenFOAM's upper-triangular
order; boundary faces that carry no physical surface go to `defaultFaces`, as gmshToFoam does.
gmsh itself is not available in this image: tests write a .msh from the repo's own tet mesher
(`write_msh`) and read it back, and read a file laid out by hand the way gmsh writes one (node ids
with gaps, shuffled nodes, CRLF, $Comments, physical points / lines, three tags, several elementary
surfaces under one physical name, mixed orientation).
"""
from __future__ import annotations

import numpy as np

from .foamfile import FoamError, PolyMesh
from .meshgen import _orient_tets, _tet_faces, build_polymesh


def read_msh(path):
    """-> (points (P,3), tets (C,4) 0-based, tri (B,3) 0-based, tri_phys (B,), phys_names {tag: (dim, name)}, tet_phys (C,))"""
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    names, points, ids, tets, tet_phys, tris, tri_phys = {}, None, None, [], [], [], []
    version = None
    while i < len(lines):
        s = lines[i].strip()
        if s == "$MeshFormat":
            version = lines[i + 1].split()[0]
            if not version.startswith("2"):
                raise FoamError(f"{path}: msh version {version} is not supported (write with -format msh2, as main.py:304-308 does)")
            if lines[i + 1].split()[1] != "0":
                raise FoamError(f"{path}: binary msh files are not supported")
            i += 3
        elif s == "$PhysicalNames":
            n = int(lines[i + 1])
            for k in range(n):
                d, tag, nm = lines[i + 2 + k].split(None, 2)
                names[int(tag)] = (int(d), nm.strip().strip('"'))
            i += n + 3
        elif s == "$Nodes":
            n = int(lines[i + 1])
            a = np.array(" ".join(lines[i + 2 : i + 2 + n]).split(), dtype=np.float64).reshape(n, 4)
            ids = a[:, 0].astype(np.int64)
            points = a[:, 1:4].copy()
            i += n + 3
        elif s == "$Elements":
            n = int(lines[i + 1])
            for k in range(n):
                t = lines[i + 2 + k].split()
                etype, ntags = int(t[1]), int(t[2])
                phys = int(t[3]) if ntags > 0 else 0
                nodes = t[3 + ntags :]
                if etype == 4:
                    tets.append(nodes)
                    tet_phys.append(phys)
                elif etype == 2:
                    tris.append(nodes)
                    tri_phys.append(phys)
                elif etype in (1, 15):
                    pass  # lines / points carry nothing for the polyMesh
                else:
                    raise FoamError(f"{path}: element type {etype} is not supported (tetrahedral meshes only)")
            i += n + 3
        else:
            i += 1
    if points is None or not tets:
        raise FoamError(f"{path}: no nodes or no tetrahedra found")
    # node ids -> 0-based contiguous
    lut = np.full(int(ids.max()) + 1, -1, dtype=np.int64)
    lut[ids] = np.arange(ids.size)
    tets = lut[np.array(tets, dtype=np.int64)]
    tris = lut[np.array(tris, dtype=np.int64)] if tris else np.zeros((0, 3), dtype=np.int64)
    return points, tets, tris, np.array(tri_phys, dtype=np.int64), names, np.array(tet_phys, dtype=np.int64)


def msh_to_polymesh(path):
    points, tets, tris, tri_phys, names, tet_phys = read_msh(path)
    tets = _orient_tets(points, tets)
    faces4, fcell = _tet_faces(tets)
    nP = points.shape[0]
    # physical surfaces in order of first appearance (gmshToFoam: one patch per physical surface)
    surf_tags = [t for t in dict.fromkeys(tri_phys.tolist())]
    patch_names = [names.get(t, (2, f"patch{t}"))[1] for t in surf_tags]
    tag_index = {t: k for k, t in enumerate(surf_tags)}
    s = np.sort(tris, axis=1)
    key = (s[:, 0] * (nP + 1) + s[:, 1]) * (nP + 1) + s[:, 2]
    order = np.argsort(key)
    ks, kp = key[order], np.array([tag_index[t] for t in tri_phys.tolist()], dtype=np.int64)[order]
    n_named = len(patch_names)
    state = {"default": False}

    def classify(fc, fn, bf):
        b = np.sort(bf[:, :3], axis=1)
        kk = (b[:, 0] * (nP + 1) + b[:, 1]) * (nP + 1) + b[:, 2]
        pos = np.minimum(np.searchsorted(ks, kk), max(ks.size - 1, 0))
        hit = ks[pos] == kk if ks.size else np.zeros(kk.size, dtype=bool)
        pid = np.where(hit, kp[pos] if ks.size else 0, n_named)
        state["default"] = bool((~hit).any())
        return pid

    mesh = build_polymesh(points, faces4, fcell, classify, patch_names + ["defaultFaces"], ["patch"] * (n_named + 1), zone_name=None)
    if mesh.patches[-1]["nFaces"] == 0:
        mesh.patches.pop()
    zones = {}
    for tag in dict.fromkeys(tet_phys.tolist()):
        nm = names.get(tag, (3, f"zone{tag}"))[1]
        zones[nm] = np.nonzero(tet_phys == tag)[0].astype(np.int32)
    mesh.cell_zones = zones
    return mesh


def write_msh(path, mesh: PolyMesh, volume_name="internalMesh"):
    """A msh 2.2 file of a tetrahedral PolyMesh (tests: stands in for `gmsh -3 -format msh2`)."""
    off, lab = mesh.face_offsets, mesh.face_labels
    if np.any(np.diff(off) != 3):
        raise ValueError("write_msh needs a tetrahedral mesh")
    faces = lab.reshape(-1, 3)
    nC = mesh.n_cells
    # cell -> its 4 faces -> 4 distinct vertices
    cell_faces = [[] for _ in range(nC)]
    for f, c in enumerate(mesh.owner):
        cell_faces[c].append(f)
    for f, c in enumerate(mesh.neighbour):
        cell_faces[c].append(f)
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
        f.write(f"$PhysicalNames\n{len(mesh.patches) + 1}\n")
        for k, p in enumerate(mesh.patches):
            f.write(f'2 {k + 1} "{p["name"]}"\n')
        vtag = len(mesh.patches) + 1
        f.write(f'3 {vtag} "{volume_name}"\n$EndPhysicalNames\n')
        f.write(f"$Nodes\n{mesh.n_points}\n")
        for i, p in enumerate(mesh.points):
            f.write(f"{i + 1} {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
        f.write("$EndNodes\n")
        nb = sum(p["nFaces"] for p in mesh.patches)
        f.write(f"$Elements\n{nb + nC}\n")
        e = 1
        for k, p in enumerate(mesh.patches):
            for fi in range(p["startFace"], p["startFace"] + p["nFaces"]):
                a, b, c = faces[fi] + 1
                f.write(f"{e} 2 2 {k + 1} {k + 1} {a} {b} {c}\n")
                e += 1
        for c in range(nC):
            v = list(dict.fromkeys(int(x) for fi in cell_faces[c] for x in faces[fi]))
            f.write(f"{e} 4 2 {vtag} 1 {v[0] + 1} {v[1] + 1} {v[2] + 1} {v[3] + 1}\n")
            e += 1
        f.write("$EndElements\n")


def main(argv=None):
    """`gmshToFoam <file.msh> [-case DIR]` (circularSloshingTank/Makefile:73): writes
    constant/polyMesh of the case from a Gmsh 2.2 ASCII mesh."""
    import os
    import sys

    from . import foamfile as ff

    argv = list(sys.argv[1:] if argv is None else argv)
    if argv and argv[0] == "gmshToFoam":
        argv.pop(0)
    case_dir, msh = os.getcwd(), None
    while argv:
        a = argv.pop(0)
        if a == "-case":
            case_dir = argv.pop(0)
        elif a.startswith("-"):
            raise SystemExit(f"gmshToFoam (tppvof): unknown option {a}")
        else:
            msh = a
    if msh is None:
        raise SystemExit("usage: gmshToFoam <file.msh> [-case DIR]")
    try:
        mesh = msh_to_polymesh(msh if os.path.isabs(msh) else os.path.join(case_dir, msh))
        mesh.check()
        ff.write_polymesh(case_dir, mesh, binary=True)
    except Exception as e:
        print(f"--> FOAM FATAL ERROR: {e}", file=sys.stderr)
        return 1
    print(f"gmshToFoam (tppvof): {mesh.n_cells} cells, {mesh.n_faces} faces, patches " + ", ".join(f"{p['name']}({p['nFaces']})" for p in mesh.patches))
    return 0


if __name__ == "__main__":
    import sys

    sys.exit(main())
