"""Host logic: dictionaries, polyMesh / field readers+writers, mesh generators (addressing is
bit-exact integer work), time names, hard errors on unsupported keywords, and the C-ABI
library's exported symbols (no compute calls: works without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import foamfile as ff
from openfoam_tpp_b200 import meshgen as mg
from openfoam_tpp_b200 import solver as sv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("binary", [True, False])
def test_polymesh_roundtrip(tmp_path, binary):
    m = mg.cylinder_mesh(0.004, 0.0221, 4, 3, "flat", "tet")
    ff.write_polymesh(str(tmp_path), m, binary=binary)
    r = ff.read_polymesh(str(tmp_path))
    assert np.array_equal(r.owner, m.owner) and np.array_equal(r.neighbour, m.neighbour)
    assert np.array_equal(r.face_offsets, m.face_offsets) and np.array_equal(r.face_labels, m.face_labels)
    if binary:
        assert np.array_equal(r.points, m.points)
    else:
        assert np.allclose(r.points, m.points, rtol=0, atol=0)  # repr() round-trips doubles
    assert [(p["name"], p["type"], p["nFaces"], p["startFace"]) for p in r.patches] == [(p["name"], p["type"], p["nFaces"], p["startFace"]) for p in m.patches]
    assert list(r.cell_zones) == ["internalMesh"] and r.cell_zones["internalMesh"].size == m.n_cells


@pytest.mark.parametrize("gen", [
    lambda: mg.cylinder_mesh(0.004, 0.0221, 5, 3, "flat", "tet"),
    lambda: mg.cylinder_mesh(0.004, 0.0221, 5, 3, "cap", "prism"),
    lambda: mg.box_mesh(3, 4, 5, cell="tet", top_patch="atmosphere"),
    lambda: mg.sloshing_tank3d_mesh(4, 6, 6),
])
def test_mesh_addressing_invariants(gen):
    """OpenFOAM ordering: owner < neighbour, upper-triangular internal faces, contiguous
    patches; every cell closed (sum of outward Sf = 0); positive volumes."""
    m = gen()
    assert m.check()
    Cf, Sf = mg.face_geometry(m)
    C, V = mg.cell_geometry(m, Cf, Sf)
    assert V.min() > 0
    div = np.zeros((m.n_cells, 3))
    np.add.at(div, m.owner, Sf)
    np.add.at(div, m.neighbour, -Sf[: m.n_internal])
    assert np.abs(div).max() < 1e-12 * np.abs(Sf).max()
    # normals point from owner to neighbour
    d = C[m.neighbour] - C[m.owner[: m.n_internal]]
    assert np.all(np.einsum("ij,ij->i", d, Sf[: m.n_internal]) > 0)


def test_cylinder_naming_contract():
    """Patches walls/atmosphere, zone internalMesh (generate_mesh.py:29-51, dynamicMeshDict:25)."""
    m = mg.cylinder_mesh(0.1, 0.02, 4, 6)
    assert [p["name"] for p in m.patches] == ["walls", "atmosphere"]
    Cf, Sf = mg.face_geometry(m)
    atm = m.patch("atmosphere")
    sl = slice(atm["startFace"], atm["startFace"] + atm["nFaces"])
    assert np.allclose(Cf[sl, 2], 0.1) and np.all(Sf[sl, 2] > 0)
    area = Sf[sl, 2].sum()  # inscribed 24-gon of the R = 0.01 circle
    assert 0.97 * np.pi * 1e-4 < area < np.pi * 1e-4
    t = mg.sloshing_tank3d_mesh(4, 6, 6)
    assert [p["name"] for p in t.patches] == ["wall"] and list(t.cell_zones) == ["all"]


def test_slab_meshes_tile_the_cylinder():
    """z-slabs (the `simple` (1 1 N) decomposition) reproduce the whole mesh: same cells, and the
    two sides of every cut hold the same faces in the same order."""
    whole = mg.cylinder_mesh(0.02, 0.02, 3, 6)
    a = mg.cylinder_mesh(0.02, 0.02, 3, 6, k0=0, k1=3, proc=(0, None, 1))
    b = mg.cylinder_mesh(0.02, 0.02, 3, 6, k0=3, k1=6, proc=(1, 0, None))
    assert a.n_cells + b.n_cells == whole.n_cells
    pa, pb = a.patch("procBoundary0to1"), b.patch("procBoundary1to0")
    assert pa["nFaces"] == pb["nFaces"] > 0 and pa["type"] == "processor"
    Cfa, Sfa = mg.face_geometry(a)
    Cfb, Sfb = mg.face_geometry(b)
    sa = slice(pa["startFace"], pa["startFace"] + pa["nFaces"])
    sb = slice(pb["startFace"], pb["startFace"] + pb["nFaces"])
    assert np.allclose(Cfa[sa], Cfb[sb]) and np.allclose(Sfa[sa], -Sfb[sb])


def test_time_names():
    for t, s in [(0.05, "0.05"), (10.0, "10"), (20.0, "20"), (0.00119048, "0.00119048"), (1e-5, "1e-05"), (1234567.0, "1.23457e+06"), (0.1 + 0.2, "0.3")]:
        assert ff.time_name(t) == s


def test_set_fields_and_case_reader(tmp_path):
    d = str(tmp_path / "case_H0.004_D0.0221_flat_R0.005_f2.0")
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=5, n_layers=4)
    c = cs.Case(d)
    a = c.fields["alpha.water"].internal_array(c.mesh.n_cells)
    C, V = mg.cell_geometry(c.mesh)
    assert set(np.unique(a)) == {0.0, 1.0} and np.array_equal(a == 1.0, C[:, 2] <= 0.002)
    k = c.cfg
    assert (k.n_alpha_subcycles, k.n_alpha_corr, k.n_correctors, k.n_non_orth, k.c_alpha) == (3, 1, 2, 0, 1.0)
    assert (k.p_rgh.type, k.p_rgh.smoother, k.p_rgh.tolerance, k.p_rgh.rel_tol) == (1, 0, 1e-8, 0.01)
    assert (k.p_rgh_final.type, k.p_rgh_final.precond, k.p_rgh_final.n_vcycles, k.p_rgh_final.n_pre_sweeps, k.p_rgh_final.max_iter) == (0, 1, 2, 2, 20)
    assert (k.rho1, k.rho2, k.nu1, k.nu2, k.sigma) == (998.2, 1.0, 1e-6, 1.48e-5, 0.0)
    assert k.patch_bc_u == [0, 1] and k.patch_bc_alpha == [0, 1] and k.patch_bc_p == [0, 1]
    assert k.motion.shape == (1001, 7) and k.probes.shape == (2, 3)
    assert c.start_name == "0"


@pytest.mark.parametrize("path,old,new,frag", [
    ("system/fvSchemes", "Gauss vanLeerV", "Gauss upwind", "div(rhoPhi,U)"),
    ("system/fvSolution", "momentumPredictor no", "momentumPredictor yes", "momentumPredictor"),
    ("system/fvSolution", "smoother        DIC;", "smoother        symGaussSeidel;", "smoother"),
    ("constant/phaseProperties", "sigma           0", "sigma           -0.07", "sigma"),
    ("system/controlDict", "incompressibleVoF", "incompressibleFluid", "solver"),
    ("0/U", "movingWallVelocity", "slip", "slip"),
])
def test_unsupported_keywords_are_hard_errors(tmp_path, path, old, new, frag):
    """No silent defaults: the error names the file and the keyword."""
    d = str(tmp_path / "c")
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=0.1, n_rings=3, n_layers=2)
    p = os.path.join(d, path)
    txt = open(p).read()
    assert old in txt
    open(p, "w").write(txt.replace(old, new, 1))
    with pytest.raises(ff.FoamError) as e:
        cs.Case(d)
    assert frag in str(e.value) and os.path.basename(path) in str(e.value)


def _declared_symbols():
    h = open(os.path.join(ROOT, "include", "tppvof.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(tpp_[a-z_0-9]+)\s*\(", h)))


def test_product_library_exports_every_declared_symbol():
    """libtppvof.so (the sm_100a build) loads without a GPU and exports the whole C-ABI."""
    if not os.path.exists(sv.LIB_PATH):  # a fresh checkout: the library is a build product (nvcc cross-compiles without a GPU)
        import __graft_entry__

        __graft_entry__.build()
    assert os.path.exists(sv.LIB_PATH), "libtppvof.so missing after __graft_entry__.build()"
    lib = ctypes.CDLL(sv.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tppvof.h but not exported"
    lib.tpp_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.tpp_version()


def test_no_cpu_fallback_in_product_path():
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = mg.cylinder_mesh(0.004, 0.0221, 3, 2)
    cfg = cs.CaseConfig()
    cfg.patch_bc_u, cfg.patch_bc_alpha, cfg.patch_bc_p = [0, 1], [0, 1], [0, 1]
    cfg.patch_inlet_alpha, cfg.patch_p0 = [0, 0], [0, 0]
    with pytest.raises(sv.SolverError) as e:
        sv.Solver(m, cfg)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_references_the_oracle():
    """Only tests/, __graft_entry__.smoke and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "openfoam-tpp_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "voforacle" not in txt and "import oracle" not in txt and "orc_" not in txt, f


def test_gmsh_msh2_ingest_roundtrip(tmp_path):
    """msh 2.2 -> polyMesh (the gmshToFoam step, Makefile:73): a .msh written from the repo's own
    tet mesh comes back with the same cells (order kept), patch names/sizes, zone and geometry."""
    from openfoam_tpp_b200 import gmsh

    m = mg.cylinder_mesh(0.02, 0.02, 3, 4, "flat", "tet")
    p = tmp_path / "cylinder.msh"
    gmsh.write_msh(str(p), m)
    r = gmsh.msh_to_polymesh(str(p))
    assert r.check()
    assert r.n_cells == m.n_cells and r.n_faces == m.n_faces and r.n_internal == m.n_internal
    assert [(q["name"], q["type"], q["nFaces"]) for q in r.patches] == [(q["name"], "patch", q["nFaces"]) for q in m.patches]
    assert list(r.cell_zones) == ["internalMesh"] and r.cell_zones["internalMesh"].size == m.n_cells
    C0, V0 = mg.cell_geometry(m)
    C1, V1 = mg.cell_geometry(r)
    assert np.allclose(V0, V1, rtol=1e-12) and np.allclose(C0, C1, atol=1e-15)
    # same connectivity: owner/neighbour pairs as sets
    a = set(zip(m.owner[: m.n_internal].tolist(), m.neighbour.tolist()))
    b = set(zip(r.owner[: r.n_internal].tolist(), r.neighbour.tolist()))
    assert a == b
    with pytest.raises(ff.FoamError):
        (tmp_path / "bad.msh").write_text("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
        gmsh.msh_to_polymesh(str(tmp_path / "bad.msh"))


def test_device_interface_summary_matches_the_host_metric(emu_lib, tmp_path):
    _interface_check(emu_lib, tmp_path)


@pytest.mark.gpu
def test_device_interface_summary_gpu(gpu_lib, tmp_path):
    _interface_check(None, tmp_path)


def _interface_check(emu_lib, tmp_path):
    """tpp_interface: the alpha = 0.5 contour statistics computed by the library's own
    kernels equal the host restatement of the reference's metric (cell -> point average, one point
    per straddling edge) on a sloshed state, also on an internally renumbered mesh, and foamRun
    -interface writes them in the reference's interface_summary.csv layout."""
    import bench
    from openfoam_tpp_b200 import foamrun, interface, meshgen

    for mesh, renum in ((meshgen.cylinder_mesh(bench.CASE["H"], bench.CASE["D"], 5, 8, "flat", "tet"), "0"), (meshgen.shuffled(meshgen.cylinder_mesh(bench.CASE["H"], bench.CASE["D"], 5, 8, "flat", "prism"), 3), "1")):
        os.environ["TPP_RENUMBER"] = renum
        try:
            g = sv.Solver(mesh, bench.make_config(mesh), lib_path=emu_lib)
        finally:
            os.environ.pop("TPP_RENUMBER", None)
        C, _ = meshgen.cell_geometry(mesh)
        a = np.clip(0.5 + (0.104 + 0.2 * C[:, 0] - 0.1 * C[:, 1] - C[:, 2]) / 0.02, 0.0, 1.0)  # a tilted, smeared surface
        g.set("alpha", a)
        t, mx, mn, me, n = g.interface_summary()
        pts = interface.iso_points(mesh, mesh.points, interface.cell_to_point(mesh, a), 0.5)
        assert n == len(pts) > 20
        assert abs(mx - pts[:, 2].max()) < 1e-12 and abs(mn - pts[:, 2].min()) < 1e-12 and abs(me - pts[:, 2].mean()) < 1e-12
        g.close()
    d = str(tmp_path / "case_H0.004_D0.0221_flat_R0.005_f2.0")
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=0.004, n_rings=4, n_layers=4, write_interval=0.002)
    foamrun.run_case(d, lib_path=emu_lib, log=None, interface=True)
    rows = open(os.path.join(d, "postProcessing", "interface", "interface_summary.csv")).read().splitlines()
    assert rows[0] == "time,max_z,min_z,mean_z,num_points" and len(rows) == 4  # t = 0 and two write times
    ref = interface.extract_interface(d, write=False)
    for r, (t, mx, mn, me, n) in zip(rows[1:], ref):
        v = [float(x) for x in r.split(",")]
        assert abs(v[0] - t) < 1e-12 and int(v[4]) == n and abs(v[1] - mx) < 1e-9 and abs(v[2] - mn) < 1e-9 and abs(v[3] - me) < 1e-9


def test_a_diverged_run_stops_with_an_error(emu_lib, tmp_path):
    """a run whose Courant number is not finite must not carry on, write time
    directories and exit 0 - tpp_step / tpp_run_to_write return an error, the Python host raises and
    foamRun exits non-zero (what `check=True`, main.py:345, relies on)."""
    import bench
    from openfoam_tpp_b200 import foamrun

    mesh = mg.cylinder_mesh(bench.CASE["H"], bench.CASE["D"], 4, 6, "flat", "tet")
    g = sv.Solver(mesh, bench.make_config(mesh), lib_path=emu_lib)
    g.set("alpha", bench.initial_alpha(mesh))
    g.init_fields()
    g.step(2)
    phi = g.get("phi")
    phi[3] = np.nan
    g.set("phi", phi)
    with pytest.raises(sv.SolverError, match="not finite"):
        g.run_to_write(5)
    with pytest.raises(sv.SolverError):
        g.step(1)
    g.close()
    d = str(tmp_path / "case_H0.004_D0.0221_flat_R0.005_f2.0")
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=0.004, n_rings=4, n_layers=4, write_interval=0.002)
    U = ff.read_field(os.path.join(d, "0", "U"))
    bad = np.zeros((ff.read_polymesh(d).n_cells, 3))
    bad[5, 0] = np.inf
    ff.write_field(os.path.join(d, "0", "U"), ff.Field(U.cls, U.name, U.dimensions, bad, U.boundary), True, location="0")
    os.environ["TPP_LIB_PATH_FOR_TEST"] = emu_lib
    with pytest.raises(Exception):
        foamrun.run_case(d, lib_path=emu_lib, log=None)
    assert not any(v > 0 for v, _ in ff.time_dirs(d))  # nothing was written after the failure


def test_async_read_back_equals_blocking(emu_lib):
    """tpp_get_async + tpp_sync return what tpp_get returns (file order, also on a renumbered mesh)."""
    import bench
    from openfoam_tpp_b200 import abi

    mesh = mg.shuffled(mg.cylinder_mesh(bench.CASE["H"], bench.CASE["D"], 5, 8, "flat", "tet"), 2)
    g = sv.Solver(mesh, bench.make_config(mesh), lib_path=emu_lib)
    g.set("alpha", bench.initial_alpha(mesh))
    g.init_fields()
    g.step(2)
    for nm in ("alpha", "U", "phi", "Uf", "p_rgh_b"):
        a = np.zeros(g.size(nm))
        assert g.L.tpp_get_async(g.h, nm.encode(), a.ctypes.data_as(abi.c_double_p), a.size) == a.size
        assert g.L.tpp_sync(g.h) == 0
        assert np.array_equal(a, g.get(nm)), nm
    g.close()


def _mutations():
    def owner_out_of_range(m, c): m.owner[-1] = m.n_cells + 7
    def negative_neighbour(m, c): m.neighbour[3] = -2
    def neighbour_not_above_owner(m, c): m.neighbour[0] = m.owner[0]
    def point_label_out_of_range(m, c): m.face_labels[0] = m.n_points + 3
    def point_not_finite(m, c): m.points[0, 0] = np.nan
    def degenerate_face(m, c): m.face_offsets[1:] -= 1; m.face_offsets[1] = 2
    def patches_leave_a_gap(m, c): m.patches[-1]["startFace"] += 1
    def zero_delta_t(m, c): c.delta_t = 0.0
    def no_alpha_subcycle(m, c): c.n_alpha_subcycles = 0
    def negative_density(m, c): c.rho2 = -1.0
    def negative_sigma(m, c): c.sigma = -0.07
    def no_solver_iterations(m, c): c.p_rgh_final.max_iter = 0
    return [(f.__name__, f) for f in (owner_out_of_range, negative_neighbour, neighbour_not_above_owner, point_label_out_of_range, point_not_finite,
                                      degenerate_face, patches_leave_a_gap, zero_delta_t, no_alpha_subcycle, negative_density, negative_sigma, no_solver_iterations)]


@pytest.mark.parametrize("name,mutate", _mutations(), ids=[n for n, _ in _mutations()])
def test_damaged_input_is_an_error_code_not_a_crash(emu_lib, name, mutate):
    """every entry point returns 0 or a negative code with a message.  A polyMesh or a configuration
    that breaks what gmshToFoam / the dictionaries guarantee (Makefile:73; upper-triangular order, patches tiling the
    boundary faces, positive time control ...) is refused by tpp_create before anything is indexed with it."""
    import copy

    import bench

    mesh = mg.cylinder_mesh(0.004, 0.0221, 3, 2, "flat", "tet")
    cfg = bench.make_config(mesh)
    sv.Solver(mesh, cfg, lib_path=emu_lib).close()  # the undamaged pair is accepted
    m, c = copy.deepcopy(mesh), copy.deepcopy(cfg)
    mutate(m, c)
    with pytest.raises(sv.SolverError) as e:
        sv.Solver(m, c, lib_path=emu_lib)
    assert "invalid mesh / configuration" in str(e.value)


def test_null_handle_and_bad_arguments_return_codes(emu_lib):
    """A null handle, an unknown array / stage, a wrong length or an out-of-range probe cell is a negative return
    code with a message (include/tppvof.h), never a crash."""
    import ctypes as C

    import bench

    L = sv.load(emu_lib)
    buf = (C.c_double * 16)()
    for call in (lambda: L.tpp_step(None, 1), lambda: L.tpp_run_to_write(None, 1), lambda: L.tpp_info(None, buf), lambda: L.tpp_init_fields(None),
                 lambda: L.tpp_get(None, b"alpha", buf, 16), lambda: L.tpp_set(None, b"alpha", buf, 16), lambda: L.tpp_size(None, b"alpha"),
                 lambda: L.tpp_stage(None, b"courant"), lambda: L.tpp_set_delta_t(None, 1e-3), lambda: L.tpp_set_probes(None, 0, None),
                 lambda: L.tpp_find_cell(None, buf), lambda: L.tpp_sync(None), lambda: L.tpp_stats(None, 0, buf)):
        assert call() < 0
        assert b"null handle" in L.tpp_last_error()
    assert L.tpp_destroy(None) == 0  # like free(NULL)
    mesh = mg.cylinder_mesh(0.004, 0.0221, 3, 2, "flat", "tet")
    g = sv.Solver(mesh, bench.make_config(mesh), lib_path=emu_lib)
    assert L.tpp_get(g.h, b"no_such_array", buf, 16) == -1 and b"unknown array" in L.tpp_last_error()
    assert L.tpp_set(g.h, b"alpha", buf, 16) == -2 and b"size mismatch" in L.tpp_last_error()
    assert L.tpp_stage(g.h, b"no_such_stage") == -1 and b"unknown stage" in L.tpp_last_error()
    assert L.tpp_set_delta_t(g.h, 0.0) == -1
    cells = (C.c_int * 2)(0, mesh.n_cells)
    assert L.tpp_set_probes(g.h, 2, cells) == -2 and b"out of range" in L.tpp_last_error()
    g.close()


def test_gmsh_style_msh2_file_with_gaps_mixed_orientation_and_several_entities(tmp_path):
    """What a file written by gmsh itself looks like and the repo's own writer never produces (gmsh is not in this
    image, so the file is laid out by hand after generate_mesh.py:15-51's groups): node ids with gaps and in shuffled
    order (OpenCASCADE boolean operations leave holes), Windows line ends, `$Comments`, physical points / lines that
    carry nothing, three tags per element, the wall made of two elementary surfaces (side and bottom) under ONE
    physical name, triangles and tetrahedra in arbitrary orientation, the volume group listed first.  The polyMesh must
    equal the one built directly from the same tetrahedra: cell order, volumes, patch order (first appearance),
    upper-triangular faces."""
    from openfoam_tpp_b200 import gmsh

    H, D = 0.02, 0.02
    pts, tets = mg.unstructured_cylinder_tets(H, D, 0.004, seed=3, iters=20)
    direct = mg.unstructured_cylinder_mesh(H, D, 0.004, seed=3, iters=20)
    rng = np.random.default_rng(5)
    nP = len(pts)
    ids = np.sort(rng.choice(np.arange(1, 3 * nP), nP, replace=False))  # gmsh node id of point k
    node_order = rng.permutation(nP)
    flip_t = rng.random(len(tets)) < 0.5
    tt = np.where(flip_t[:, None], tets[:, [1, 0, 2, 3]], tets)
    # boundary triangles = faces seen once
    f = np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 1, 3]], tets[:, [0, 2, 3]], tets[:, [1, 2, 3]]])
    key = np.sort(f, axis=1)
    _, first, cnt = np.unique(key, axis=0, return_index=True, return_counts=True)
    tri = f[first[cnt == 1]]
    z = pts[tri][:, :, 2]
    top, bottom = np.abs(z - H).max(1) < 1e-9, np.abs(z).max(1) < 1e-9
    tri = np.where((rng.random(len(tri)) < 0.5)[:, None], tri[:, [0, 2, 1]], tri)
    rows = ['$MeshFormat', '2.2 0 8', '$EndMeshFormat', '$Comments', 'written by hand, gmsh layout', '$EndComments',
            '$PhysicalNames', '5', '0 7 "corner"', '1 8 "rim"', '2 1 "atmosphere"', '2 2 "walls"', '3 3 "internalMesh"', '$EndPhysicalNames',
            '$Nodes', str(nP)]
    rows += [f"{ids[k]} {float(pts[k, 0])!r} {float(pts[k, 1])!r} {float(pts[k, 2])!r}" for k in node_order]
    rows += ['$EndNodes', '$Elements']
    el = [f"15 2 7 1 {ids[0]}", f"1 2 8 4 {ids[tri[0, 0]]} {ids[tri[0, 1]]}"]
    # walls first (side = elementary 5, bottom = elementary 6), then the top: patch order must follow first appearance
    for k in np.nonzero(~top & ~bottom)[0]:
        el.append(f"2 3 2 5 0 {ids[tri[k, 0]]} {ids[tri[k, 1]]} {ids[tri[k, 2]]}")
    for k in np.nonzero(bottom)[0]:
        el.append(f"2 3 2 6 0 {ids[tri[k, 0]]} {ids[tri[k, 1]]} {ids[tri[k, 2]]}")
    for k in np.nonzero(top)[0]:
        el.append(f"2 2 1 4 {ids[tri[k, 0]]} {ids[tri[k, 1]]} {ids[tri[k, 2]]}")
    for t in tt:
        el.append(f"4 2 3 1 {ids[t[0]]} {ids[t[1]]} {ids[t[2]]} {ids[t[3]]}")
    rows += [str(len(el))] + [f"{10 + 3 * i} {e}" for i, e in enumerate(el)] + ['$EndElements']  # element numbers with gaps too
    p = tmp_path / "cylinder.msh"
    p.write_bytes(("\r\n".join(rows) + "\r\n").encode())
    r = gmsh.msh_to_polymesh(str(p))
    assert r.check()
    assert [(q["name"], q["type"]) for q in r.patches] == [("walls", "patch"), ("atmosphere", "patch")]
    assert [q["nFaces"] for q in r.patches] == [q["nFaces"] for q in direct.patches]
    assert (r.n_cells, r.n_faces, r.n_internal) == (direct.n_cells, direct.n_faces, direct.n_internal)
    assert list(r.cell_zones) == ["internalMesh"] and np.array_equal(r.cell_zones["internalMesh"], np.arange(r.n_cells))
    # points come back in the file's node order; cells keep the $Elements order: same cells as the direct mesh
    C0, V0 = mg.cell_geometry(direct)
    C1, V1 = mg.cell_geometry(r)
    assert (V1 > 0).all() and np.allclose(V0, V1, rtol=1e-11) and np.allclose(C0, C1, atol=1e-14)
    assert np.array_equal(r.owner[: r.n_internal], direct.owner[: direct.n_internal]) and np.array_equal(r.neighbour, direct.neighbour)
    assert (r.neighbour > r.owner[: r.n_internal]).all()


def test_field_files_round_trip_property(tmp_path):
    """volScalar / volVector / surfaceScalar fields, uniform and nonuniform, ascii and binary, through
    write_field / read_field: binary payloads come back bit for bit whatever bytes they contain (a double whose
    bytes spell `);` or `}` or `//` must not end the list), ascii ones to writePrecision 6
    (system/controlDict:35-37); empty patches and zero-size lists survive."""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from hypothesis.extra import numpy as hnp

    tricky = np.frombuffer((b");\n}\n//;(" + b"/*(;)*/{" + b"\n)\n;\n//\n" + b"FoamFile")[:32], dtype="<f8")  # 4 doubles of syntax-looking bytes
    finite = st.floats(-1e12, 1e12, allow_nan=False, width=64)
    counter = [0]

    @settings(max_examples=60, deadline=None)
    @given(cls=st.sampled_from(["volScalarField", "volVectorField", "surfaceScalarField"]), binary=st.booleans(), n=st.integers(0, 40),
           uniform=st.booleans(), data=st.data())
    def check(cls, binary, n, uniform, data):
        nc = 3 if cls == "volVectorField" else 1
        shape = (n, 3) if nc == 3 else (n,)
        if uniform:
            internal = np.array(data.draw(st.lists(finite, min_size=3, max_size=3))) if nc == 3 else data.draw(finite)
        else:
            internal = data.draw(hnp.arrays(np.float64, shape, elements=finite))
            if binary and n >= 4:
                internal.reshape(-1)[:4] = tricky
        nb = data.draw(st.integers(0, 6))
        bval = data.draw(hnp.arrays(np.float64, (nb, 3) if nc == 3 else (nb,), elements=finite))
        boundary = {"walls": {"type": "zeroGradient"}, "atmosphere": {"type": "inletOutlet", "inletValue": "uniform 0", "value": bval},
                    "empty_one": {"type": "calculated", "value": np.zeros((0, 3) if nc == 3 else (0,))}}
        counter[0] += 1
        p = str(tmp_path / f"f{counter[0]}" / "0" / "fld")
        ff.write_field(p, ff.Field(cls, "fld", "[0 1 -1 0 0 0 0]", internal, boundary), binary=binary, location="0")
        r = ff.read_field(p)
        assert r.cls == cls and list(r.boundary) == ["walls", "atmosphere", "empty_one"]
        assert r.boundary["walls"]["type"] == "zeroGradient" and r.boundary["atmosphere"]["type"] == "inletOutlet"
        same = (lambda a, b: np.asarray(a, dtype="<f8").tobytes() == np.asarray(b, dtype="<f8").tobytes()) if binary else \
               (lambda a, b: np.allclose(np.asarray(a), np.asarray(b), rtol=6e-6, atol=0))
        if uniform:
            assert np.allclose(np.asarray(r.internal), np.asarray(internal), rtol=6e-6, atol=0)  # `uniform` values are text in both formats
        else:
            assert np.asarray(r.internal).shape == shape and same(r.internal, internal)
        assert np.asarray(r.boundary["atmosphere"]["value"]).reshape(-1).size == bval.size and same(np.asarray(r.boundary["atmosphere"]["value"]).reshape(bval.shape), bval)
        assert np.asarray(r.boundary["empty_one"]["value"]).size == 0

    check()
