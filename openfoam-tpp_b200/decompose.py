"""decomposePar / reconstructPar equivalents for the multi-GPU path.

The reference decomposes a case with OpenFOAM's `decomposePar` before `mpirun -np N foamRun
-parallel` and merges the result with `reconstructPar`
(/root/reference/circularSloshingTank/Makefile:77-82, system/decomposeParDict). This module writes
and reads the same on-disk layout - `processorN/constant/polyMesh/{points,faces,owner,neighbour,
boundary,cellProcAddressing,faceProcAddressing,pointProcAddressing,boundaryProcAddressing}` and
`processorN/<time>/<field>` - so a rank of the GPU solver starts from exactly what a rank of
`foamRun -parallel` would read, and existing decomposed cases can be resumed.

What is restated from OpenFOAM-13 ([OF13-MEM]: written from the algorithm as remembered, the
source is not available here; the integer outputs are self-checked by tests/test_decompose.py,
not pinned against an OpenFOAM run):

* `simple` (simpleGeomDecomp): cell centres are rotated by the small skew `rotDelta(delta)`,
  sorted along x, y, z in turn, and cut into n.x / n.y / n.z groups of equal count (the first
  `size % n` groups get one more); processor = ix + n.x*iy + n.x*n.y*iz.
* `hierarchical`: the same equal-count cuts applied recursively in the given `order`
  (OpenFOAM's version bisects on coordinate values with a tolerance; on ties the two can differ).
* `scotch` is NOT reproduced (needs the library): a request for it falls back to `hierarchical`
  with the factorisation of numberOfSubdomains closest to a cube, and says so.
* domainDecomposition: cells of a processor in ascending global index; its faces are (1) the
  internal faces with both cells on it, in global order, (2) every original patch in order
  (kept, possibly empty) with the faces whose owner it holds, (3) one `processor` patch per
  neighbouring processor in ascending processor number, faces in ascending global face index -
  the same order on both sides, which is what the solver's halo exchange relies on; on the
  side that holds the global *neighbour* cell the face is reversed (`face::reverseFace`: first
  point kept) and its faceProcAddressing entry is negative. faceProcAddressing stores global
  face + 1. Points of a processor in ascending global index.
"""
from __future__ import annotations

import os

import numpy as np

from . import foamfile as ff
from . import meshgen
from .foamfile import FoamError, PolyMesh


# ---- cell -> processor --------------------------------------------------------------------------
def rot_delta(delta):
    d = 1.0 - 0.5 * delta * delta
    d2, a = d * d, delta
    a2 = a * a
    return np.array([[d2, -a * d, a], [a * d - a2 * d, a * a2 + d2, -2 * a * d], [a * d2 + a2, a * d - a2 * d, d2 - a2]])


def _assign_groups(size, n):
    """simpleGeomDecomp::assignToProcessorGroup: group index of every sorted position."""
    jump = size // n
    fst = size - jump * n
    g = np.empty(size, dtype=np.int64)
    g[: fst * (jump + 1)] = np.repeat(np.arange(fst), jump + 1)
    g[fst * (jump + 1) :] = np.repeat(np.arange(fst, n), jump)
    return g


def partition_simple(centres, n, delta=0.001):
    n = tuple(int(x) for x in n)
    pts = centres @ rot_delta(delta).T
    proc = np.zeros(len(pts), dtype=np.int64)
    mult = 1
    for comp in range(3):
        order = np.argsort(pts[:, comp], kind="stable")
        proc[order] += mult * _assign_groups(len(pts), n[comp])
        mult *= n[comp]
    return proc.astype(np.int32)


def partition_hierarchical(centres, n, order="xyz", delta=0.001):
    n = tuple(int(x) for x in n)
    pts = centres @ rot_delta(delta).T
    comps = ["xyz".index(c) for c in order]
    mults = {0: 1, 1: n[0], 2: n[0] * n[1]}
    proc = np.zeros(len(pts), dtype=np.int64)

    def rec(idx, level):
        if level == 3:
            return
        comp = comps[level]
        o = idx[np.argsort(pts[idx, comp], kind="stable")]
        g = _assign_groups(len(o), n[comp])
        proc[o] += mults[comp] * g
        for k in range(n[comp]):
            rec(o[g == k], level + 1)

    rec(np.arange(len(pts)), 0)
    return proc.astype(np.int32)


def _cube_factors(n):
    best = (n, 1, 1)
    for a in range(1, n + 1):
        if n % a:
            continue
        for b in range(1, n // a + 1):
            if (n // a) % b:
                continue
            c = n // a // b
            t = tuple(sorted((a, b, c), reverse=True))
            if max(t) - min(t) < max(best) - min(best):
                best = t
    return best


def read_decompose_dict(case_dir):
    """system/decomposeParDict -> (nProcs, method, n, order, delta)."""
    path = os.path.join(case_dir, "system", "decomposeParDict")
    d = ff.read_dict(path)
    nproc = int(ff.to_float(ff.lookup(d, "numberOfSubdomains", path)))
    method = str(ff.lookup(d, "method", path))
    n, order, delta = None, "xyz", 0.001
    co = d.get(f"{method}Coeffs", d.get("coeffs"))
    if method in ("simple", "hierarchical"):
        if co is None:
            raise FoamError(f"{path}: {method}Coeffs missing")
        n = tuple(int(x) for x in ff.to_vector(ff.lookup(co, "n", path)))
        if n[0] * n[1] * n[2] != nproc:
            raise FoamError(f"{path}: n {n} does not multiply to numberOfSubdomains {nproc}")
        delta = ff.to_float(co["delta"]) if "delta" in co else 0.001
        order = str(co.get("order", "xyz"))
    elif method == "scotch":
        n = _cube_factors(nproc)
    else:
        raise FoamError(f"{path}: decomposition method '{method}' is not supported (simple, hierarchical, scotch->hierarchical)")
    return nproc, method, n, order, delta


def cell_partition(mesh, nproc, method, n, order="xyz", delta=0.001, log=None):
    C, _ = meshgen.cell_geometry(mesh)
    if method == "simple":
        return partition_simple(C, n, delta)
    if method == "scotch" and log:
        print(f"decomposePar (tppvof): scotch is not available, using hierarchical {n} xyz (not bit-identical to scotch)", file=log)
    return partition_hierarchical(C, n, order, delta)


# ---- mesh decomposition ---------------------------------------------------------------------------
class ProcMesh:
    def __init__(self, mesh, cell_addr, face_addr, point_addr, boundary_addr):
        self.mesh = mesh
        self.cell_addr = cell_addr          # local cell -> global cell
        self.face_addr = face_addr          # local face -> +-(global face + 1); negative = reversed
        self.point_addr = point_addr        # local point -> global point
        self.boundary_addr = boundary_addr  # local patch -> global patch (-1 for processor patches)


def decompose_mesh(mesh: PolyMesh, cell_proc, nproc=None):
    cell_proc = np.asarray(cell_proc, dtype=np.int64)
    nproc = int(cell_proc.max()) + 1 if nproc is None else nproc
    nI = mesh.n_internal
    own, nei = mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64)
    po = cell_proc[own]                 # processor of every face's owner
    pn = cell_proc[nei]                 # ... of every internal face's neighbour
    off, lab = mesh.face_offsets.astype(np.int64), mesh.face_labels.astype(np.int64)
    parts = []
    for me in range(nproc):
        cells = np.flatnonzero(cell_proc == me)
        g2l = np.full(mesh.n_cells, -1, dtype=np.int64)
        g2l[cells] = np.arange(len(cells))
        internal = np.flatnonzero((po[:nI] == me) & (pn == me))
        faces = [internal]
        flip = [np.zeros(len(internal), dtype=bool)]
        patches, baddr = [], []
        start = len(internal)
        for gi, p in enumerate(mesh.patches):
            if p["type"] == "processor":
                raise FoamError("decompose_mesh expects an undecomposed mesh")
            f = np.arange(p["startFace"], p["startFace"] + p["nFaces"])
            f = f[po[f] == me]
            q = {k: v for k, v in p.items()}
            q["nFaces"], q["startFace"] = len(f), start
            patches.append(q)
            baddr.append(gi)
            faces.append(f)
            flip.append(np.zeros(len(f), dtype=bool))
            start += len(f)
        cut = np.flatnonzero((po[:nI] == me) != (pn == me))  # internal faces between me and another processor
        mine_is_owner = po[cut] == me
        other = np.where(mine_is_owner, pn[cut], po[cut])
        for nb in np.unique(other):
            sel = other == nb
            f = cut[sel]                     # ascending global face index
            patches.append({"name": f"procBoundary{me}to{int(nb)}", "type": "processor", "nFaces": len(f), "startFace": start, "myProcNo": me, "neighbProcNo": int(nb)})
            baddr.append(-1)
            faces.append(f)
            flip.append(~mine_is_owner[sel])
            start += len(f)
        faces = np.concatenate(faces)
        flip = np.concatenate(flip)
        # local owner / neighbour
        lown = np.where(flip, g2l[nei[np.minimum(faces, nI - 1)]] if nI else -1, g2l[own[faces]])
        lnei = g2l[nei[internal]]
        # points and faces
        cnt = off[faces + 1] - off[faces]
        foff = np.concatenate([[0], np.cumsum(cnt)])
        idx = np.repeat(off[faces], cnt) + (np.arange(foff[-1]) - np.repeat(foff[:-1], cnt))
        flab = lab[idx]
        # reversed faces keep their first point: (p0, pn-1, ..., p1)
        if flip.any():
            pos = np.arange(foff[-1]) - np.repeat(foff[:-1], cnt)
            rev = np.repeat(flip, cnt)
            c_ = np.repeat(cnt, cnt)
            src = np.where(rev & (pos > 0), c_ - pos, pos)
            flab = lab[np.repeat(off[faces], cnt) + src]
        pts = np.unique(flab)
        p2l = np.full(mesh.n_points, -1, dtype=np.int64)
        p2l[pts] = np.arange(len(pts))
        zones = {}
        for nm, z in (mesh.cell_zones or {}).items():
            zl = g2l[np.asarray(z, dtype=np.int64)]
            zones[nm] = np.sort(zl[zl >= 0]).astype(np.int32)
        pm = PolyMesh(mesh.points[pts], foff, p2l[flab], lown, lnei, patches, zones)
        parts.append(ProcMesh(pm, cells.astype(np.int32), np.where(flip, -(faces + 1), faces + 1).astype(np.int32), pts.astype(np.int32), np.asarray(baddr, dtype=np.int32)))
    return parts


# ---- files ----------------------------------------------------------------------------------------
def _write_labels(path, name, arr, location, note=None):
    with open(path, "wb") as f:
        f.write(ff._hdr("labelIOList", name, location, "binary", note).encode())
        f.write(f"\n{arr.size}\n(".encode())
        f.write(np.asarray(arr).astype("<i4").tobytes())
        f.write(b")\n")
        f.write(ff.END.encode())


def _read_labels(path):
    a, _ = ff._read_label_list(path)
    return a


def _is_list(v, nc):
    """a per-face / per-cell list (not a uniform value) of a field with nc components"""
    return isinstance(v, np.ndarray) and v.ndim == (1 if nc == 1 else 2)


def _split_field(fld: ff.Field, part: ProcMesh, whole: PolyMesh):
    """A vol / surface field restricted to one processor.  A processor patch carries the value
    decomposePar gives it: the neighbour-side cell value (vol fields) or the face value itself
    (surface fields; a flux - surfaceScalarField - changes sign on a reversed face)."""
    pm = part.mesh
    nc = ff._NCOMP[ff._CLASS_TYPE[fld.cls]]
    surface = fld.cls.startswith("surface")
    oriented = fld.cls == "surfaceScalarField"
    fa = np.abs(part.face_addr.astype(np.int64)) - 1
    flipped = part.face_addr < 0
    uniform = not _is_list(np.asarray(fld.internal), nc)
    full = fld.internal_array(whole.n_internal if surface else whole.n_cells)
    internal = fld.internal if uniform else (full[fa[: pm.n_internal]] if surface else full[part.cell_addr])
    boundary = {}
    for p in pm.patches:
        sl = slice(p["startFace"], p["startFace"] + p["nFaces"])
        if p["type"] == "processor":
            if surface:
                v = full[fa[sl]]
                if oriented:
                    v = v * np.where(flipped[sl], -1.0, 1.0)
            else:
                other = np.where(flipped[sl], whole.owner[fa[sl]], whole.neighbour[np.minimum(fa[sl], max(whole.n_internal - 1, 0))])
                v = full[other]
            boundary[p["name"]] = {"type": "processor", "value": v}
            continue
        src = dict(fld.boundary[p["name"]])
        gp = whole.patch(p["name"])
        for k, v in list(src.items()):
            if _is_list(v, nc) and v.shape[0] == gp["nFaces"]:
                src[k] = v[fa[sl] - gp["startFace"]]
        boundary[p["name"]] = src
    return ff.Field(fld.cls, fld.name, fld.dimensions, internal, boundary)


def decompose_par(case_dir, time_name=None, binary=True, log=None):
    """decomposePar: system/decomposeParDict, constant/polyMesh and the start fields ->
    processor0..N-1.  Returns the list of ProcMesh."""
    from .case import latest_time

    mesh = ff.read_polymesh(case_dir)
    nproc, method, n, order, delta = read_decompose_dict(case_dir)
    cell_proc = cell_partition(mesh, nproc, method, n, order, delta, log)
    parts = decompose_mesh(mesh, cell_proc, nproc)
    if time_name is None:
        time_name = latest_time(case_dir)[1]
    tdir = os.path.join(case_dir, time_name)
    fields = []
    for nm in sorted(os.listdir(tdir)):
        fp = os.path.join(tdir, nm)
        if os.path.isfile(fp) and _is_field_file(fp):
            fields.append(ff.read_field(fp))  # a corrupt field is an error, not a silently dropped file
    for k, part in enumerate(parts):
        pd = os.path.join(case_dir, f"processor{k}")
        ff.write_polymesh(pd, part.mesh, binary)
        md = os.path.join(pd, "constant", "polyMesh")
        _write_labels(os.path.join(md, "cellProcAddressing"), "cellProcAddressing", part.cell_addr, "constant/polyMesh")
        _write_labels(os.path.join(md, "faceProcAddressing"), "faceProcAddressing", part.face_addr, "constant/polyMesh")
        _write_labels(os.path.join(md, "pointProcAddressing"), "pointProcAddressing", part.point_addr, "constant/polyMesh")
        _write_labels(os.path.join(md, "boundaryProcAddressing"), "boundaryProcAddressing", part.boundary_addr, "constant/polyMesh")
        os.makedirs(os.path.join(pd, time_name), exist_ok=True)
        for fld in fields:
            ff.write_field(os.path.join(pd, time_name, fld.name), _split_field(fld, part, mesh), binary, location=time_name)
        up = os.path.join(tdir, "uniform", "time")
        if os.path.exists(up):
            os.makedirs(os.path.join(pd, time_name, "uniform"), exist_ok=True)
            with open(up, "rb") as src, open(os.path.join(pd, time_name, "uniform", "time"), "wb") as dst:
                dst.write(src.read())
    if log:
        sizes = [p.mesh.n_cells for p in parts]
        print(f"decomposePar (tppvof): {nproc} processors, method {method} {n}, cells {min(sizes)}..{max(sizes)}, processor faces {sum(q['nFaces'] for p in parts for q in p.mesh.patches if q['type'] == 'processor') // 2}", file=log)
    return parts


_FIELD_CLASSES = ("volScalarField", "volVectorField", "volTensorField", "volSymmTensorField", "surfaceScalarField", "surfaceVectorField")


def _is_field_file(path):
    """True when the FoamFile header names a geometric field class; other files of a time directory
    (dictionaries, logs) are left alone by decomposePar / reconstructPar."""
    try:
        with open(path, "rb") as f:
            hdr = ff.parse_header(f.read(4096), path)[0]
    except Exception:
        return False
    return hdr.get("class") in _FIELD_CLASSES


def processor_dirs(case_dir):
    k = 0
    out = []
    while os.path.isdir(os.path.join(case_dir, f"processor{k}")):
        out.append(os.path.join(case_dir, f"processor{k}"))
        k += 1
    return out


def reconstruct_par(case_dir, times=None, binary=True, log=None, with_zero=False, new_times=False):
    """reconstructPar: merge processorN/<time>/<field> into <time>/<field> through the
    *ProcAddressing files.  vol fields and surface fields of the solver's output set.
    Like OpenFOAM's utility, time 0 is skipped unless `with_zero` (-withZero) or named in
    `times` (so the user's 0/ files are not overwritten by flattened copies), and `new_times`
    (-newTimes) skips the times that already exist at the case root."""
    mesh = ff.read_polymesh(case_dir)
    pdirs = processor_dirs(case_dir)
    if not pdirs:
        raise FoamError(f"{case_dir}: no processor directories")
    addr = []
    for pd in pdirs:
        md = os.path.join(pd, "constant", "polyMesh")
        addr.append((ff.read_polymesh(pd), _read_labels(os.path.join(md, "cellProcAddressing")), _read_labels(os.path.join(md, "faceProcAddressing")), _read_labels(os.path.join(md, "boundaryProcAddressing"))))
    if times:
        tnames = list(times)
    else:
        tnames = [nm for v, nm in ff.time_dirs(pdirs[0]) if with_zero or v != 0]
        if new_times:
            have = {nm for _, nm in ff.time_dirs(case_dir)}
            tnames = [nm for nm in tnames if nm not in have]
    done = []
    for tn in tnames:
        names = [nm for nm in sorted(os.listdir(os.path.join(pdirs[0], tn))) if os.path.isfile(os.path.join(pdirs[0], tn, nm)) and _is_field_file(os.path.join(pdirs[0], tn, nm))]
        os.makedirs(os.path.join(case_dir, tn), exist_ok=True)
        for nm in names:
            parts = [ff.read_field(os.path.join(pd, tn, nm)) for pd in pdirs]
            f0 = parts[0]
            surface = f0.cls.startswith("surface")
            scalar = f0.cls.endswith("ScalarField")
            ncomp = 1 if scalar else (9 if f0.cls.endswith("TensorField") else 3)
            shape = (mesh.n_internal if surface else mesh.n_cells,) + (() if scalar else (ncomp,))
            glob = np.zeros(shape)
            bvals = {p["name"]: {} for p in mesh.patches}
            for (pm, ca, fa, ba), fld in zip(addr, parts):
                gfa = np.abs(fa.astype(np.int64)) - 1
                if surface:
                    glob[gfa[: pm.n_internal]] = fld.internal_array(pm.n_internal)
                else:
                    glob[ca] = fld.internal_array(pm.n_cells)
                for p in pm.patches:
                    e = fld.boundary.get(p["name"], {})
                    sl = slice(p["startFace"], p["startFace"] + p["nFaces"])
                    if p["type"] == "processor":
                        # a cut face is an internal face of the whole mesh: its value comes from the
                        # side that kept the global orientation (positive faceProcAddressing)
                        v = e.get("value")
                        if surface and _is_list(v, ncomp) and v.shape[0] == p["nFaces"] and p["nFaces"]:
                            own_side = fa[sl] > 0
                            glob[gfa[sl][own_side]] = v[own_side]
                        continue
                    gp = mesh.patch(p["name"])
                    dst = bvals[p["name"]]
                    for k, v in e.items():
                        if _is_list(v, ncomp) and v.shape[0] == p["nFaces"]:
                            if k not in dst or not _is_list(dst[k], ncomp):
                                dst[k] = np.zeros((gp["nFaces"],) + v.shape[1:])
                            if p["nFaces"]:
                                dst[k][gfa[sl] - gp["startFace"]] = v
                        elif k not in dst:
                            dst[k] = v
            ff.write_field(os.path.join(case_dir, tn, nm), ff.Field(f0.cls, f0.name, f0.dimensions, glob, bvals), binary, location=tn)
        # moved points of a dynamic mesh (<time>/polyMesh/points) through pointProcAddressing
        if os.path.exists(os.path.join(pdirs[0], tn, "polyMesh", "points")):
            pts = np.zeros_like(mesh.points)
            for pd in pdirs:
                pa = _read_labels(os.path.join(pd, "constant", "polyMesh", "pointProcAddressing"))
                pts[pa] = ff.read_points(os.path.join(pd, tn, "polyMesh", "points"))
            os.makedirs(os.path.join(case_dir, tn, "polyMesh"), exist_ok=True)
            ff.write_points(os.path.join(case_dir, tn, "polyMesh", "points"), pts, binary, f"{tn}/polyMesh")
        up = os.path.join(pdirs[0], tn, "uniform", "time")
        if os.path.exists(up):
            os.makedirs(os.path.join(case_dir, tn, "uniform"), exist_ok=True)
            with open(up, "rb") as src, open(os.path.join(case_dir, tn, "uniform", "time"), "wb") as dst:
                dst.write(src.read())
        done.append(tn)
    if log:
        print(f"reconstructPar (tppvof): {len(done)} time(s) from {len(pdirs)} processors", file=log)
    return done


def main(argv=None):
    """`python -m openfoam_tpp_b200.decompose decomposePar|reconstructPar [-case DIR]` - drop-in for
    the two OpenFOAM utilities the reference's Makefile calls around `foamRun -parallel`."""
    import sys

    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("decomposePar", "reconstructPar"):
        raise SystemExit("usage: decompose decomposePar|reconstructPar [-case DIR] [-force] [-latestTime]")
    tool, case_dir, latest, with_zero, new_times = argv.pop(0), os.getcwd(), False, False, False
    while argv:
        a = argv.pop(0)
        if a == "-case":
            case_dir = argv.pop(0)
        elif a == "-latestTime":
            latest = True
        elif a == "-withZero":
            with_zero = True
        elif a == "-newTimes":
            new_times = True
        elif a == "-force":
            pass
        else:
            raise SystemExit(f"{tool} (tppvof): unknown option {a}")
    try:
        if tool == "decomposePar":
            decompose_par(case_dir, log=sys.stdout)
        else:
            times = None
            if latest:
                times = [ff.time_dirs(processor_dirs(case_dir)[0])[-1][1]]
            reconstruct_par(case_dir, times, log=sys.stdout, with_zero=with_zero, new_times=new_times)
    except Exception as e:
        print(f"--> FOAM FATAL ERROR: {e}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    import sys

    sys.exit(main())
