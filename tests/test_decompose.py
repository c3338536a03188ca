"""decomposePar / reconstructPar equivalents (SURVEY.md §8e partitioning, §8f rank 2): integer
addressing self-checks (bit-exact work), field round trips, and the `foamRun -parallel` driver on a
decomposed case (2 ranks, gloo, host emulation of the kernels) against the serial run."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import decompose as dc
from openfoam_tpp_b200 import foamfile as ff
from openfoam_tpp_b200 import meshgen as mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_groups_matches_simpleGeomDecomp():
    # 10 points into 3 groups: jump 3, one group of 4 first (assignToProcessorGroup)
    assert dc._assign_groups(10, 3).tolist() == [0, 0, 0, 0, 1, 1, 1, 2, 2, 2]
    assert dc._assign_groups(9, 3).tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2]
    assert dc._assign_groups(5, 1).tolist() == [0] * 5


def test_simple_partition_on_a_box():
    mesh = mg.box_mesh(6, 4, 4)
    C, _ = mg.cell_geometry(mesh)
    p = dc.partition_simple(C, (3, 2, 1))
    assert np.bincount(p).tolist() == [16] * 6
    # processor = ix + nx * iy: x thirds and y halves of the box (delta's skew cannot reorder a regular lattice)
    ix = np.minimum((C[:, 0] * 3).astype(int), 2)
    iy = np.minimum((C[:, 1] * 2).astype(int), 1)
    assert np.array_equal(p, ix + 3 * iy)
    ph = dc.partition_hierarchical(C, (3, 2, 1), "xyz")
    assert np.array_equal(ph, p)


@pytest.mark.parametrize("cell,n", [("tet", (1, 1, 3)), ("hex", (2, 2, 1))])
def test_decompose_mesh_addressing(cell, n):
    mesh = mg.cylinder_mesh(0.05, 0.04, 4, 6, "flat", "tet") if cell == "tet" else mg.box_mesh(5, 4, 3)
    C, _ = mg.cell_geometry(mesh)
    proc = dc.partition_simple(C, n)
    parts = dc.decompose_mesh(mesh, proc)
    nI = mesh.n_internal
    assert sum(p.mesh.n_cells for p in parts) == mesh.n_cells
    seen_int = np.zeros(mesh.n_faces, dtype=int)
    seen_pos = np.zeros(mesh.n_faces, dtype=int)
    seen_neg = np.zeros(mesh.n_faces, dtype=int)
    Sf = {}
    for k, part in enumerate(parts):
        pm = part.mesh
        pm.check()  # upper-triangular order, contiguous patches
        assert np.array_equal(np.sort(part.cell_addr), np.flatnonzero(proc == k))
        assert np.all(np.diff(part.cell_addr) > 0) and np.all(np.diff(part.point_addr) > 0)
        assert np.array_equal(pm.points, mesh.points[part.point_addr])
        fa = np.abs(part.face_addr.astype(np.int64)) - 1
        assert np.all(part.face_addr[: pm.n_internal] > 0) and np.all(np.diff(fa[: pm.n_internal]) > 0)
        seen_int[fa[: pm.n_internal]] += 1
        # every face keeps its point set; a reversed face keeps its first point and reverses the rest
        for lf in range(pm.n_faces):
            g = mesh.face_labels[mesh.face_offsets[fa[lf]] : mesh.face_offsets[fa[lf] + 1]]
            l = part.point_addr[pm.face_labels[pm.face_offsets[lf] : pm.face_offsets[lf + 1]]]
            if part.face_addr[lf] > 0:
                assert np.array_equal(l, g)
            else:
                assert l[0] == g[0] and np.array_equal(l[1:], g[1:][::-1])
        # owner / neighbour map back to the global cells
        gown = part.cell_addr[pm.owner]
        flipped = part.face_addr < 0
        assert np.array_equal(gown[~flipped], mesh.owner[fa[~flipped]])
        assert np.array_equal(gown[flipped], mesh.neighbour[fa[flipped]])
        assert np.array_equal(part.cell_addr[pm.neighbour], mesh.neighbour[fa[: pm.n_internal]])
        names = [q["name"] for q in pm.patches]
        assert names[: len(mesh.patches)] == [q["name"] for q in mesh.patches]  # original patches kept, in order
        nbs = [q["neighbProcNo"] for q in pm.patches if q["type"] == "processor"]
        assert nbs == sorted(nbs) and all(q["myProcNo"] == k for q in pm.patches if q["type"] == "processor")
        _, sf = mg.face_geometry(pm)
        for q in pm.patches:
            sl = slice(q["startFace"], q["startFace"] + q["nFaces"])
            if q["type"] == "processor":
                assert np.all(fa[sl] < nI) and np.all(np.diff(fa[sl]) > 0)  # ascending global face order
                seen_pos[fa[sl][part.face_addr[sl] > 0]] += 1
                seen_neg[fa[sl][part.face_addr[sl] < 0]] += 1
                Sf[(k, q["neighbProcNo"])] = (fa[sl], sf[sl])
            else:
                gp = mesh.patch(q["name"])
                assert np.all((fa[sl] >= gp["startFace"]) & (fa[sl] < gp["startFace"] + gp["nFaces"]))
                seen_int[fa[sl]] += 1
    cut = (seen_pos == 1) & (seen_neg == 1)
    assert np.all((seen_int == 1) ^ cut) and not np.any(seen_pos + seen_neg > 2)
    # both sides of an interface list the same faces in the same order, with opposite area vectors
    for (a, b), (faces, sf) in Sf.items():
        faces2, sf2 = Sf[(b, a)]
        assert np.array_equal(faces, faces2)
        assert np.allclose(sf, -sf2, rtol=0, atol=1e-15)


def _fields(case_dir, tn):
    return {nm: ff.read_field(os.path.join(case_dir, tn, nm)) for nm in sorted(os.listdir(os.path.join(case_dir, tn))) if os.path.isfile(os.path.join(case_dir, tn, nm))}


def test_decompose_reconstruct_roundtrip(tmp_path):
    d = str(tmp_path / "case")
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=5, n_layers=6)
    mesh = ff.read_polymesh(d)
    # a flux-like and a vector surface field with distinct values on every face
    rng = np.random.default_rng(3)
    nI, nF = mesh.n_internal, mesh.n_faces
    phi = rng.standard_normal(nF)
    calc = lambda v, nc: {p["name"]: {"type": "calculated", "value": (v[p["startFace"] : p["startFace"] + p["nFaces"]])} for p in mesh.patches}
    ff.write_field(os.path.join(d, "0", "phi"), ff.Field("surfaceScalarField", "phi", "[0 3 -1 0 0 0 0]", phi[:nI], calc(phi, 1)), True, location="0")
    Uf = rng.standard_normal((nF, 3))
    ff.write_field(os.path.join(d, "0", "Uf"), ff.Field("surfaceVectorField", "Uf", "[0 1 -1 0 0 0 0]", Uf[:nI], calc(Uf, 3)), True, location="0")
    with open(os.path.join(d, "system", "decomposeParDict"), "w") as f:
        f.write(ff._hdr("dictionary", "decomposeParDict", "system") + "numberOfSubdomains 4;\nmethod hierarchical;\nhierarchicalCoeffs { n (2 1 2); order xyz; delta 0.001; }\n" + ff.END)
    before = _fields(d, "0")
    parts = dc.decompose_par(d)
    assert len(parts) == 4 and len(dc.processor_dirs(d)) == 4
    # the files read back as written (addressing is integer work: exact)
    for k, part in enumerate(parts):
        pd = os.path.join(d, f"processor{k}")
        pm = ff.read_polymesh(pd)
        assert np.array_equal(pm.owner, part.mesh.owner) and np.array_equal(pm.neighbour, part.mesh.neighbour)
        assert np.array_equal(pm.face_labels, part.mesh.face_labels) and np.array_equal(pm.points, part.mesh.points)
        assert [q.get("neighbProcNo") for q in pm.patches] == [q.get("neighbProcNo") for q in part.mesh.patches]
        assert np.array_equal(dc._read_labels(os.path.join(pd, "constant/polyMesh/faceProcAddressing")), part.face_addr)
        # a flux on a reversed processor face changes sign, a vector does not
        f = ff.read_field(os.path.join(pd, "0", "phi"))
        for q in pm.patches:
            if q["type"] == "processor":
                sl = slice(q["startFace"], q["startFace"] + q["nFaces"])
                g = np.abs(part.face_addr[sl].astype(np.int64)) - 1
                assert np.array_equal(f.boundary[q["name"]]["value"], phi[g] * np.sign(part.face_addr[sl]))
    for nm in os.listdir(os.path.join(d, "0")):
        if os.path.isfile(os.path.join(d, "0", nm)):
            os.remove(os.path.join(d, "0", nm))
    dc.reconstruct_par(d, ["0"])
    after = _fields(d, "0")
    assert set(after) == set(before)
    for nm, b in before.items():
        a = after[nm]
        n = mesh.n_internal if b.cls.startswith("surface") else mesh.n_cells
        assert np.array_equal(a.internal_array(n), b.internal_array(n)), nm
        for p in mesh.patches:
            vb, va = b.boundary[p["name"]].get("value"), a.boundary[p["name"]].get("value")
            if isinstance(vb, np.ndarray) and vb.ndim >= 1 and vb.shape[0] == p["nFaces"] and p["nFaces"] != 3:
                assert np.array_equal(va, vb), (nm, p["name"])
            assert a.boundary[p["name"]]["type"] == b.boundary[p["name"]]["type"]


WORKER = """
import os, sys
sys.path.insert(0, {root!r})
from openfoam_tpp_b200 import foamrun
out = foamrun.run_case({case!r}, lib_path={lib!r}, max_steps={steps}, parallel=True, log=None)
sys.stdout.write('RANK%sOK %d %d\\n' % (os.environ['RANK'], out['steps'], out['cells'])); sys.stdout.flush()
import torch.distributed as dist
dist.destroy_process_group()
"""


def test_foamrun_parallel_matches_serial(tmp_path, emu_lib):
    """decomposePar -> `foamRun -parallel` on 2 ranks -> reconstructPar == the serial run."""
    from openfoam_tpp_b200 import foamrun

    steps = 6
    serial, par = str(tmp_path / "serial"), str(tmp_path / "par")
    for d in (serial, par):
        cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=6, n_layers=8, write_interval=0.002, p_final_max_iter=400)
        for fn in ("fvSolution",):
            p = os.path.join(d, "system", fn)
            s = open(p).read().replace("tolerance       1e-08;", "tolerance       1e-13;").replace("tolerance       2e-09;", "tolerance       1e-13;").replace("relTol          0.01;", "relTol          0;")
            open(p, "w").write(s)
    with open(os.path.join(par, "system", "decomposeParDict"), "w") as f:
        f.write(ff._hdr("dictionary", "decomposeParDict", "system") + "numberOfSubdomains 2;\nmethod simple;\nsimpleCoeffs { n (1 1 2); delta 0.001; }\n" + ff.END)
    out = foamrun.run_case(serial, lib_path=emu_lib, max_steps=steps, log=None)
    dc.decompose_par(par)
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(WORKER.format(root=ROOT, case=par, lib=emu_lib, steps=steps)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29641", str(script)],
                       capture_output=True, text=True, timeout=900)
    o = r.stdout + r.stderr
    assert r.returncode == 0 and "RANK0OK" in o and "RANK1OK" in o, o[-3000:]
    times = [nm for _, nm in ff.time_dirs(os.path.join(par, "processor0")) if nm != "0"]
    assert times and times == [nm for _, nm in ff.time_dirs(serial) if nm != "0"][: len(times)]
    dc.reconstruct_par(par, times)
    mesh = ff.read_polymesh(serial)
    for tn in times:
        a, b = _fields(par, tn), _fields(serial, tn)
        for nm, tol in (("alpha.water", 1e-9), ("U", 1e-7), ("p_rgh", 1e-7), ("phi", 1e-7), ("Uf", 1e-7)):
            n = mesh.n_internal if b[nm].cls.startswith("surface") else mesh.n_cells
            x, y = a[nm].internal_array(n), b[nm].internal_array(n)
            assert np.abs(x - y).max() <= tol * max(np.abs(y).max(), 1e-300), (tn, nm, np.abs(x - y).max())
        # the moved mesh of that time, merged through pointProcAddressing
        pa, pb = ff.read_points(os.path.join(par, tn, "polyMesh", "points")), ff.read_points(os.path.join(serial, tn, "polyMesh", "points"))
        assert np.abs(pa - pb).max() <= 1e-15
    assert out["steps"] == steps
    # probes: written once, by the master, at the case root; same rows as the serial run
    pr_s = open(os.path.join(serial, "postProcessing", "probes", "0", "p")).read().splitlines()
    pr_p = open(os.path.join(par, "postProcessing", "probes", "0", "p")).read().splitlines()
    assert len(pr_p) == len(pr_s) and pr_p[:3] == pr_s[:3]
    for a, b in zip(pr_p[3:], pr_s[3:]):
        va, vb = [float(x) for x in a.split()], [float(x) for x in b.split()]
        assert va[0] == vb[0]
        assert all(abs(x - y) <= 1e-5 * max(abs(y), 1.0) for x, y in zip(va[1:], vb[1:])), (a, b)


def test_closed_tank_parallel_matches_serial(tmp_path, emu_lib):
    """The tutorial's own parallel set-up in small: sloshingTank3D6DoF (closed, one wall patch, p_rgh
    pinned by pRefPoint / pRefValue, rotating 6-DoF motion) decomposed `hierarchical` as in
    sloshingTank3D6DoF/system/decomposeParDict:17-29 - the reference cell lives on one rank only and
    the others learn its pressure - gives the serial run's fields."""
    from openfoam_tpp_b200 import foamrun

    steps = 4
    serial, par = str(tmp_path / "serial"), str(tmp_path / "par")
    for d in (serial, par):
        cs.setup_tutorial_case(d, nx=6, ny=10, nz=8, end_time=1.0, p_final_max_iter=400)
        p = os.path.join(d, "system", "fvSolution")
        s = open(p).read().replace("tolerance       1e-08;", "tolerance       1e-13;").replace("tolerance       2e-09;", "tolerance       1e-13;").replace("relTol          0.01;", "relTol          0;")
        open(p, "w").write(s)
        p = os.path.join(d, "system", "controlDict")
        s = open(p).read().replace("writeInterval   0.05;", "writeInterval   0.02;")
        open(p, "w").write(s)
    with open(os.path.join(par, "system", "decomposeParDict"), "w") as f:
        f.write(ff._hdr("dictionary", "decomposeParDict", "system") + "numberOfSubdomains 3;\nmethod hierarchical;\nhierarchicalCoeffs { n (1 3 1); order xyz; delta 0.001; }\n" + ff.END)
    foamrun.run_case(serial, lib_path=emu_lib, max_steps=steps, log=None)
    dc.decompose_par(par)
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(WORKER.format(root=ROOT, case=par, lib=emu_lib, steps=steps)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1", "--master-port", "29647", str(script)],
                       capture_output=True, text=True, timeout=900)
    o = r.stdout + r.stderr
    assert r.returncode == 0 and all(f"RANK{k}OK" in o for k in range(3)), o[-3000:]
    times = [nm for _, nm in ff.time_dirs(os.path.join(par, "processor0")) if nm != "0"]
    assert times and times == [nm for _, nm in ff.time_dirs(serial) if nm != "0"][: len(times)]
    dc.reconstruct_par(par, times)
    mesh = ff.read_polymesh(serial)
    for tn in times:
        a, b = _fields(par, tn), _fields(serial, tn)
        for nm, tol in (("alpha.water", 1e-9), ("U", 1e-7), ("p_rgh", 1e-7), ("p", 1e-7)):
            x, y = a[nm].internal_array(mesh.n_cells), b[nm].internal_array(mesh.n_cells)
            assert np.abs(x - y).max() <= tol * max(np.abs(y).max(), 1e-300), (tn, nm, np.abs(x - y).max())
