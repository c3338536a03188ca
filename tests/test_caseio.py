"""The library's own case reader / writer (tpp_open, tpp_write_time, tpp_run_case: include/tppvof.h,
csrc/tpp_caseio.h) against the Python host's (case.Case, foamrun.run_case).  Both stand in for
`foamRun` started in a case directory (/root/reference/circularSloshingTank/Makefile:85,98;
main.py:333-348) and must be interchangeable: the same solver state from the same files, the same
files from the same state.  Host logic only - runs on the emulation build of the kernels."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import foamfile as ff
from openfoam_tpp_b200 import foamrun
from openfoam_tpp_b200 import solver as sv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATE = ("alpha", "U", "p_rgh", "p", "rho", "phi", "Uf", "alpha_b", "U_b", "p_rgh_b", "rho_b", "points", "V", "Sf")


def _setup(d, **kw):
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=5, n_layers=5, write_interval=0.003, **kw)


def _set_entry(path, key, value):
    s = open(path).read()
    s, n = re.subn(rf"(\n\s*{re.escape(key)}\s+)[^;]+;", lambda m: m.group(1) + value + ";", s, count=1)
    assert n == 1, (path, key)
    open(path, "w").write(s)


def _python_solver(d, lib):
    c = cs.Case(d)
    s = sv.Solver(c.mesh, c.cfg, lib_path=lib)
    s.load_case_fields(c)
    return c, s


def _same_state(a, b, names=STATE):
    for nm in names:
        x, y = a.get(nm), b.get(nm)
        assert x.shape == y.shape, nm
        assert np.array_equal(x, y), (nm, np.abs(x - y).max())
    ia, ib = a.info(), b.info()
    for k in ("t", "dt", "step", "Co", "alphaCo"):
        assert ia[k] == ib[k], (k, ia[k], ib[k])


def test_open_gives_the_python_hosts_solver_state_and_steps_alike(tmp_path, emu_lib):
    d = str(tmp_path / "case")
    _setup(d)
    c, a = _python_solver(d, emu_lib)
    b = sv.Solver.open(d, lib_path=emu_lib)
    assert b.case_query("n_cells") == c.mesh.n_cells and b.case_query("n_faces") == c.mesh.n_faces
    assert b.case_query("n_points") == c.mesh.n_points and b.case_query("n_patches") == len(c.mesh.patches)
    assert b.case_query("start_time") == "0" and b.case_query("n_probes") == len(c.cfg.probes)
    for nm in ("owner", "neighbour", "cf", "cn", "layout"):
        assert np.array_equal(a.get_int(nm), b.get_int(nm)), nm
    _same_state(a, b)
    a.step(3)
    b.step(3)
    _same_state(a, b)  # configuration (schemes, controls, motion table, solver settings) read alike
    a.close()
    b.close()


@pytest.mark.parametrize("binary", [True, False])
def test_write_time_matches_the_python_writer(tmp_path, emu_lib, binary):
    """same state -> same files: binary files byte for byte; ascii files value for value"""
    d = str(tmp_path / "case")
    _setup(d)
    _set_entry(os.path.join(d, "system", "controlDict"), "writeFormat", "binary" if binary else "ascii")
    _set_entry(os.path.join(d, "system", "controlDict"), "writePrecision", "12")
    c, a = _python_solver(d, emu_lib)
    b = sv.Solver.open(d, lib_path=emu_lib)
    assert b.case_query("write_binary") == int(binary)
    a.step(2)
    b.step(2)
    name = ff.time_name(a.info()["t"], c.cfg.time_precision)
    assert b.case_query("time") == name
    b.write_time()
    written = os.path.join(d, name)
    kept = str(tmp_path / "from_library")
    shutil.move(written, kept)
    foamrun.write_time(c, a, name, c.cfg.write_binary, c.cfg.write_precision)
    for nm in ("alpha.water", "U", "p_rgh", "p", "rho", "phi", "Uf", os.path.join("polyMesh", "points")):
        x, y = open(os.path.join(kept, nm), "rb").read(), open(os.path.join(written, nm), "rb").read()
        if binary or nm != os.path.join("polyMesh", "points"):  # ascii points: repr() vs %.17g, same values
            assert x == y, nm
        else:
            assert np.array_equal(ff.read_points(os.path.join(kept, nm)), ff.read_points(os.path.join(written, nm)))
    ta, tb = ff.read_dict(os.path.join(kept, "uniform", "time")), ff.read_dict(os.path.join(written, "uniform", "time"))
    for k in ("value", "deltaT", "deltaT0", "index"):
        assert ff.to_float(ta[k]) == ff.to_float(tb[k]), k
    assert ta["name"] == tb["name"]
    a.close()
    b.close()


@pytest.mark.parametrize("binary", [True, False])
def test_run_case_writes_what_foamrun_writes_and_resumes(tmp_path, emu_lib, binary):
    """tpp_run_case == foamrun.run_case (time directories, probes log); a second tpp_open resumes from
    the newest complete directory and lands where the uninterrupted Python run does."""
    a, b = str(tmp_path / "python"), str(tmp_path / "library")
    for d in (a, b):
        _setup(d)
        _set_entry(os.path.join(d, "system", "controlDict"), "writeFormat", "binary" if binary else "ascii")
        _set_entry(os.path.join(d, "system", "controlDict"), "writePrecision", "17")
        _set_entry(os.path.join(d, "system", "controlDict"), "endTime", "0.006")
    out = foamrun.run_case(a, lib_path=emu_lib, log=None)
    _set_entry(os.path.join(b, "system", "controlDict"), "endTime", "0.003")
    s = sv.Solver.open(b, lib_path=emu_lib)
    n1 = s.run_case()
    s.close()
    os.makedirs(os.path.join(b, "0.0045", "uniform"))  # a directory a killed run left half-written
    open(os.path.join(b, "0.0045", "alpha.water"), "w").write("garbage")
    open(os.path.join(b, "0.0045", "phi"), "w").write("garbage")
    _set_entry(os.path.join(b, "system", "controlDict"), "endTime", "0.006")
    s = sv.Solver.open(b, lib_path=emu_lib)
    assert s.case_query("start_time") == "0.003"
    n2 = s.run_case()
    s.close()
    assert n1 + n2 == out["steps"] and n1 > 0 and n2 > 0
    mesh = ff.read_polymesh(a)
    for t in ("0.003", "0.006"):
        for nm in ("alpha.water", "U", "p_rgh", "p", "rho", "phi", "Uf"):
            fa, fb = ff.read_field(os.path.join(a, t, nm)), ff.read_field(os.path.join(b, t, nm))
            n = mesh.n_internal if fa.cls.startswith("surface") else mesh.n_cells
            x, y = fa.internal_array(n), fb.internal_array(n)
            if t == "0.003":
                assert np.array_equal(x, y), (t, nm)
            else:  # through a restart: the tolerance tests/test_restart.py gives the Python host
                assert np.abs(x - y).max() <= 1e-9 * max(np.abs(x).max(), 1e-300), (t, nm)
            assert list(fa.boundary) == list(fb.boundary)
            for pn in fa.boundary:
                assert {k: v for k, v in fa.boundary[pn].items() if k != "value"} == {k: v for k, v in fb.boundary[pn].items() if k != "value"}
    pa = open(os.path.join(a, "postProcessing", "probes", "0", "p")).read().splitlines()
    pb = open(os.path.join(b, "postProcessing", "probes", "0", "p")).read().splitlines()
    assert pa[: len(pb)] == pb and len(pb) == n1 + 1 + (len(pa) - out["steps"] - 1)  # header + row 0 + one row per step
    assert os.path.exists(os.path.join(b, "postProcessing", "probes", "0.003", "p"))


def test_reads_what_the_python_writer_wrote_in_ascii(tmp_path, emu_lib):
    """ascii polyMesh (faceList `3(a b c)` entries) and ascii nonuniform start fields"""
    d = str(tmp_path / "case")
    _setup(d)
    mesh = ff.read_polymesh(d)
    ff.write_polymesh(d, mesh, binary=False)
    c = cs.Case(d)
    for nm in ("alpha.water", "U", "p_rgh"):
        f = c.fields[nm]
        f.internal = f.internal_array(mesh.n_cells) + (0.0 if nm != "p_rgh" else np.linspace(0, 1, mesh.n_cells))
        ff.write_field(os.path.join(d, "0", nm), f, binary=False, precision=17, location="0")
    c, a = _python_solver(d, emu_lib)
    b = sv.Solver.open(d, lib_path=emu_lib)
    _same_state(a, b)
    a.close()
    b.close()


def test_tutorial_hex_tank_opens_alike(tmp_path, emu_lib):
    """sloshingTank3D6DoF layout: hexahedra, rotation in the 6DoF table, pRefPoint"""
    d = str(tmp_path / "tank")
    cs.setup_tutorial_case(d, nx=4, ny=6, nz=5, end_time=0.1)
    c, a = _python_solver(d, emu_lib)
    b = sv.Solver.open(d, lib_path=emu_lib)
    a.step(2)
    b.step(2)
    _same_state(a, b)
    a.close()
    b.close()


@pytest.mark.parametrize(
    "path,key,value,needle",
    [
        ("system/fvSchemes", "div(rhoPhi,U)", "Gauss upwind", "div(rhoPhi,U)"),
        ("system/fvSolution", "momentumPredictor", "yes", "momentumPredictor"),
        ("system/controlDict", "writeControl", "timeStep", "writeControl"),
        ("constant/momentumTransport", "simulationType", "RAS", "simulationType"),
        ("system/fvSolution", "correctPhi", "yes", "correctPhi"),
        ("constant/physicalProperties.water", "viscosityModel", "CrossPowerLaw", "viscosityModel"),
    ],
)
def test_unsupported_keywords_are_errors_that_name_the_file(tmp_path, emu_lib, path, key, value, needle):
    d = str(tmp_path / "case")
    _setup(d)
    _set_entry(os.path.join(d, path), key, value)
    with pytest.raises(ff.FoamError):
        cs.Case(d)  # the Python host refuses the same case
    with pytest.raises(sv.SolverError) as e:
        sv.Solver.open(d, lib_path=emu_lib)
    assert needle in str(e.value) and os.path.basename(path) in str(e.value) and "(-4)" in str(e.value)


def test_broken_case_directories_are_error_codes(tmp_path, emu_lib):
    d = str(tmp_path / "case")
    with pytest.raises(sv.SolverError, match="not a case directory"):
        sv.Solver.open(d, lib_path=emu_lib)
    _setup(d)
    os.rename(os.path.join(d, "0", "U"), os.path.join(d, "0", "U.away"))
    with pytest.raises(sv.SolverError, match="0/U"):
        sv.Solver.open(d, lib_path=emu_lib)
    os.rename(os.path.join(d, "0", "U.away"), os.path.join(d, "0", "U"))
    raw = open(os.path.join(d, "constant", "polyMesh", "owner"), "rb").read()
    open(os.path.join(d, "constant", "polyMesh", "owner"), "wb").write(raw[: len(raw) // 2])
    with pytest.raises(sv.SolverError, match="owner"):
        sv.Solver.open(d, lib_path=emu_lib)
    with pytest.raises(sv.SolverError, match="processor3"):
        sv.Solver.open(d, lib_path=emu_lib, processor=3)


def test_c_host_runs_a_case_directory(tmp_path, emu_lib):
    """tools/tpp_foamrun.c: a host with no Python and no FoamFile code of its own"""
    exe = str(tmp_path / "tpp_foamrun")
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "tpp_foamrun.c"), "-o", exe, "-ldl"], check=True)
    d = str(tmp_path / "case")
    _setup(d)
    _set_entry(os.path.join(d, "system", "controlDict"), "endTime", "0.003")
    r = subprocess.run([exe, emu_lib, "-case", d], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Time = 0.003" in r.stdout and r.stdout.rstrip().splitlines()[-1].startswith("End")
    assert cs.latest_time(d)[1] == "0.003"
    _set_entry(os.path.join(d, "system", "fvSchemes"), "div(rhoPhi,U)", "Gauss upwind")
    r = subprocess.run([exe, emu_lib, "-case", d], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "FOAM FATAL ERROR" in r.stderr and "fvSchemes" in r.stderr


def test_reads_openfoams_own_binary_field(emu_lib):
    """golden G1: 0/alpha.water as OpenFOAM-13's setFields wrote it (7 766 cells, 3 886 of them water)"""
    import json

    here = os.path.dirname(os.path.abspath(__file__))
    ref = json.load(open(os.path.join(here, "golden", "golden.json")))["G1_alpha"]["case_H0.004_D0.0221_flat_R0.005_f2.0"]
    a, uniform = sv.read_field_file(os.path.join(here, "golden", "alpha.water.G1"), lib_path=emu_lib)
    assert not uniform and a.size == ref["n"] == 7766 and float(a.sum()) == ref["sum"]
    assert np.array_equal(a, ff.read_field(os.path.join(here, "golden", "alpha.water.G1")).internal)
    with pytest.raises(sv.SolverError, match="cannot open"):
        sv.read_field_file(os.path.join(here, "golden", "absent"), lib_path=emu_lib)


REFERENCE_TEMPLATE = "/root/reference/circularSloshingTank"


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TEMPLATE), reason="the reference tree exists in the build container only")
def test_the_references_own_dictionaries_and_start_fields(tmp_path, emu_lib):
    """the reference's template case as it stands (circularSloshingTank/{system,constant,0}) on a mesh
    of this repo's generator: both hosts read it, agree on the configuration and step alike"""
    d = str(tmp_path / "case")
    _setup(d)
    for sub in ("system", "constant", "0"):
        for nm in os.listdir(os.path.join(REFERENCE_TEMPLATE, sub)):
            src = os.path.join(REFERENCE_TEMPLATE, sub, nm)
            if os.path.isfile(src) and nm != "setFieldsDict":
                shutil.copyfile(src, os.path.join(d, sub, nm))
    _set_entry(os.path.join(d, "system", "controlDict"), "writeInterval", "0.003")
    c, a = _python_solver(d, emu_lib)
    b = sv.Solver.open(d, lib_path=emu_lib)
    assert c.cfg.p_rgh_final.precond == 1 and c.cfg.n_alpha_subcycles == 3  # fvSolution:19-23,50-66
    a.step(2)
    b.step(2)
    _same_state(a, b)
    a.close()
    b.close()


PAR_WORKER = """
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from openfoam_tpp_b200 import foamrun, solver as sv
rank = int(os.environ['RANK'])
foamrun.run_case({py_case!r}, lib_path={lib!r}, parallel=True, log=None)      # the Python host, rank by rank
s = sv.Solver.open({lib_case!r}, lib_path={lib!r}, processor=rank)           # the library's reader on processor<rank>/
s.comm_init_callbacks()
s.case_start()
n = s.run_case()
sys.stdout.write('RANK%dOK %d %s\\\\n' % (rank, n, s.case_query('dir'))); sys.stdout.flush()
s.close()
dist.destroy_process_group()
"""


def test_processor_shares_open_like_the_python_host(tmp_path, emu_lib):
    """`foamRun -parallel` on processor<k>/ (Makefile:78): tpp_open(processor=k) -> comm -> tpp_case_start
    -> tpp_run_case writes the processor time directories the Python host writes"""
    import sys
    import textwrap

    from openfoam_tpp_b200 import decompose as dc

    a, b = str(tmp_path / "python"), str(tmp_path / "library")
    for d in (a, b):
        _setup(d)
        _set_entry(os.path.join(d, "system", "controlDict"), "endTime", "0.003")
        with open(os.path.join(d, "system", "decomposeParDict"), "w") as f:
            f.write(ff._hdr("dictionary", "decomposeParDict", "system") + "numberOfSubdomains 2;\nmethod simple;\nsimpleCoeffs { n (2 1 1); delta 0.001; }\n" + ff.END)
        dc.decompose_par(d)
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(PAR_WORKER.format(root=ROOT, py_case=a, lib_case=b, lib=emu_lib)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29647", str(script)],
                       capture_output=True, text=True, timeout=900)
    o = r.stdout + r.stderr
    assert r.returncode == 0 and "RANK0OK" in o and "RANK1OK" in o, o[-3000:]
    for k in (0, 1):
        pa, pb = os.path.join(a, f"processor{k}"), os.path.join(b, f"processor{k}")
        assert cs.latest_time(pa)[1] == cs.latest_time(pb)[1] == "0.003"
        for nm in ("alpha.water", "U", "p_rgh", "p", "rho", "phi", "Uf"):
            assert open(os.path.join(pa, "0.003", nm), "rb").read() == open(os.path.join(pb, "0.003", nm), "rb").read(), (k, nm)


def test_run_case_interface_rows_match_foamrun(tmp_path, emu_lib):
    """`foamRun -interface`: interface_summary.csv (main.py:751-780) written in situ by both hosts"""
    a, b = str(tmp_path / "python"), str(tmp_path / "library")
    for d in (a, b):
        _setup(d)
        _set_entry(os.path.join(d, "system", "controlDict"), "endTime", "0.006")
    foamrun.run_case(a, lib_path=emu_lib, log=None, interface=True)
    s = sv.Solver.open(b, lib_path=emu_lib)
    s.run_case(interface=True)
    s.close()
    fa = open(os.path.join(a, "postProcessing", "interface", "interface_summary.csv")).read().splitlines()
    fb = open(os.path.join(b, "postProcessing", "interface", "interface_summary.csv")).read().splitlines()
    assert fa[0] == fb[0] == "time,max_z,min_z,mean_z,num_points" and len(fa) == len(fb) == 4  # t = 0, 0.003, 0.006
    for x, y in zip(fa[1:], fb[1:]):
        assert [float(v) for v in x.split(",")] == [float(v) for v in y.split(",")]


def test_product_library_has_no_cpu_path_behind_tpp_open(tmp_path):
    """without a CUDA device tpp_open on the PRODUCT library is an error, never a silent CPU run"""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    if not os.path.exists(sv.LIB_PATH):
        pytest.skip("libtppvof.so not built")
    d = str(tmp_path / "case")
    _setup(d)
    with pytest.raises(sv.SolverError, match="no usable CUDA device"):
        sv.Solver.open(d)


def test_library_reader_agrees_with_the_python_reader_property(tmp_path, emu_lib):
    """random field files (scalar / vector, vol / surface, uniform / nonuniform, ascii / binary, payload bytes
    that spell `);`, `}` or `//`, empty patches): tpp_read_field returns the Python reader's values bit for bit"""
    from hypothesis import given, settings
    from hypothesis import strategies as st
    from hypothesis.extra import numpy as hnp

    tricky = np.frombuffer((b");\n}\n//;(" + b"/*(;)*/{" + b"\n)\n;\n//\n" + b"FoamFile")[:32], dtype="<f8")
    finite = st.floats(-1e12, 1e12, allow_nan=False, width=64)
    counter = [0]

    @settings(max_examples=60, deadline=None)
    @given(cls=st.sampled_from(["volScalarField", "volVectorField", "surfaceScalarField", "surfaceVectorField"]), binary=st.booleans(), n=st.integers(0, 40),
           uniform=st.booleans(), prec=st.sampled_from([6, 12, 17]), data=st.data())
    def check(cls, binary, n, uniform, prec, data):
        nc = 3 if "Vector" in cls else 1
        shape = (n, 3) if nc == 3 else (n,)
        if uniform:
            internal = np.array(data.draw(st.lists(finite, min_size=3, max_size=3))) if nc == 3 else data.draw(finite)
        else:
            internal = data.draw(hnp.arrays(np.float64, shape, elements=finite))
            if binary and n >= 4:
                internal.reshape(-1)[:4] = tricky
        nb = data.draw(st.integers(0, 6))
        bval = data.draw(hnp.arrays(np.float64, (nb, 3) if nc == 3 else (nb,), elements=finite))
        if binary and bval.size >= 4:
            bval.reshape(-1)[:4] = tricky
        boundary = {"walls": {"type": "zeroGradient"}, "atmosphere": {"type": "inletOutlet", "inletValue": "uniform 0", "value": bval},
                    "empty_one": {"type": "calculated", "value": np.zeros((0, 3) if nc == 3 else (0,))}}
        counter[0] += 1
        p = str(tmp_path / f"f{counter[0]}" / "0" / "fld")
        ff.write_field(p, ff.Field(cls, "fld", "[0 1 -1 0 0 0 0]", internal, boundary), binary=binary, precision=prec, location="0")
        want = ff.read_field(p).internal
        got, uni = sv.read_field_file(p, lib_path=emu_lib)
        assert uni == uniform
        assert np.asarray(want, dtype="<f8").tobytes() == np.asarray(got, dtype="<f8").tobytes()
        if not uniform:
            assert got.shape == shape

    check()


def test_boundary_file_as_gmshtofoam_writes_it_in_a_binary_case(tmp_path, emu_lib):
    """polyMesh/boundary of a `writeFormat binary` case: binary header, text body, `inGroups List<word> 1(wall);`"""
    d = str(tmp_path / "case")
    _setup(d)
    p = os.path.join(d, "constant", "polyMesh", "boundary")
    s = open(p).read()
    s = s.replace("format      ascii;", 'format      binary;\n    arch        "LSB;label=32;scalar=64";')
    s, n = re.subn(r"(\n\s*type\s+patch;)", r"\1\n        physicalType    patch;\n        inGroups        List<word> 1(wall);", s)
    assert n == 2
    open(p, "w").write(s)
    c, a = _python_solver(d, emu_lib)
    b = sv.Solver.open(d, lib_path=emu_lib)
    assert [q["type"] for q in c.mesh.patches] == ["patch", "patch"] and b.case_query("n_patches") == 2
    _same_state(a, b)
    a.close()
    b.close()


def test_case_handles_run_concurrently_from_host_threads(tmp_path, emu_lib):
    """the sweep's execution model (main.py:504-534: independent cases) at the case-directory level: one host
    thread and one handle per case, all running at once, each ending where it ends when run alone"""
    import threading

    dirs = []
    for k in range(4):
        d = str(tmp_path / f"case{k}")
        cs.setup_case(d, H=0.004, D=0.0221, R=0.004 + 0.001 * k, freq=2.0, duration=1.0, n_rings=5, n_layers=5, write_interval=0.003, end_time=0.006)
        dirs.append(d)
    alone = str(tmp_path / "alone")
    shutil.copytree(dirs[2], alone)
    s = sv.Solver.open(alone, lib_path=emu_lib)
    n_alone = s.run_case()
    s.close()
    out, errs = {}, []

    def work(d):
        try:
            h = sv.Solver.open(d, lib_path=emu_lib)
            out[d] = h.run_case()
            h.close()
        except Exception as e:  # noqa: BLE001
            errs.append((d, e))

    ts = [threading.Thread(target=work, args=(d,)) for d in dirs]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert out[dirs[2]] == n_alone
    for nm in ("alpha.water", "U", "p_rgh", "phi"):
        assert open(os.path.join(alone, "0.006", nm), "rb").read() == open(os.path.join(dirs[2], "0.006", nm), "rb").read(), nm
    # a failing case reports its own message to its own thread (thread-local tpp_last_error)
    _set_entry(os.path.join(dirs[0], "system", "fvSchemes"), "div(rhoPhi,U)", "Gauss upwind")
    shutil.rmtree(os.path.join(dirs[1], "constant", "polyMesh"))
    msgs = {}

    def fail(d):
        try:
            sv.Solver.open(d, lib_path=emu_lib).close()
        except sv.SolverError as e:
            msgs[d] = str(e)

    ts = [threading.Thread(target=fail, args=(d,)) for d in dirs[:2]]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert "fvSchemes" in msgs[dirs[0]] and "polyMesh" in msgs[dirs[1]]


def test_start_fields_of_the_wrong_size_are_refused(tmp_path, emu_lib):
    d = str(tmp_path / "case")
    _setup(d)
    c = cs.Case(d)
    f = c.fields["alpha.water"]
    f.internal = f.internal_array(c.mesh.n_cells)[:-1]
    ff.write_field(os.path.join(d, "0", "alpha.water"), f, binary=True, location="0")
    with pytest.raises(sv.SolverError, match=r"alpha.water:internalField: holds \d+ values, the mesh needs \d+"):
        sv.Solver.open(d, lib_path=emu_lib)
    f = c.fields["U"]
    f.boundary["walls"]["value"] = np.zeros((3, 3))
    ff.write_field(os.path.join(d, "0", "alpha.water"), c.fields["alpha.water"].__class__(f.cls.replace("Vector", "Scalar"), "alpha.water", "[0 0 0 0 0 0 0]", 0.0, c.fields["alpha.water"].boundary), binary=True, location="0")
    ff.write_field(os.path.join(d, "0", "U"), f, binary=True, location="0")
    with pytest.raises(sv.SolverError, match=r"walls.value: holds 3 values"):
        sv.Solver.open(d, lib_path=emu_lib)


def test_face_offsets_that_overrun_the_labels_are_refused_by_both_hosts(tmp_path, emu_lib):
    """found by mutation fuzzing under AddressSanitizer: the array ABI has no length for face_labels, so a
    damaged faceCompactList must be stopped by whoever read the file"""
    d = str(tmp_path / "case")
    _setup(d)
    mesh = ff.read_polymesh(d)
    mesh.face_offsets = mesh.face_offsets.copy()
    mesh.face_offsets[-1] += 7
    ff.write_polymesh(d, mesh, binary=True)
    with pytest.raises(sv.SolverError, match="faces: the offsets do not span"):
        sv.Solver.open(d, lib_path=emu_lib)
    c = cs.Case(d)
    with pytest.raises(sv.SolverError, match="invalid mesh / configuration: polyMesh faces: the offsets do not span"):
        sv.Solver(c.mesh, c.cfg, lib_path=emu_lib)


def test_the_binding_integration_md_gives_for_main_py(tmp_path, emu_lib):
    """the ctypes-only `run_case_local` a maintainer pastes into the reference's
    main.py (333-348) - extracted from the document and run as it stands"""
    import subprocess as sp

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    start = text.index("import ctypes, subprocess\n")
    code = text[start : text.index("```", start)].replace("/path/to/openfoam-tpp_b200/libtppvof.so", emu_lib)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    d = str(tmp_path / "case")
    _setup(d)
    _set_entry(os.path.join(d, "system", "controlDict"), "endTime", "0.003")
    ns["run_case_local"](d)
    assert cs.latest_time(d)[1] == "0.003"
    _set_entry(os.path.join(d, "system", "controlDict"), "endTime", "0.006")
    ns["run_case_local"](d, n_cpus=4)  # resume
    assert cs.latest_time(d)[1] == "0.006"
    _set_entry(os.path.join(d, "system", "fvSolution"), "momentumPredictor", "yes")
    with pytest.raises(sp.CalledProcessError) as e:
        ns["run_case_local"](d)
    assert "momentumPredictor" in e.value.stderr
