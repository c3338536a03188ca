"""tpp_open / tpp_run_case on the product library (sm_100a): the library's own case reader gives the
CUDA solver the Python host's start state, and a C host runs a case directory on the GPU to the
time directories the Python `foamRun` writes.  (Sorted last on purpose: the stage- and step-parity
tests against the oracle come first.)"""
import os
import subprocess

import numpy as np
import pytest

from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import foamfile as ff
from openfoam_tpp_b200 import foamrun
from openfoam_tpp_b200 import solver as sv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(d, end_time):
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=8, n_layers=6, write_interval=0.003, end_time=end_time)


@pytest.mark.gpu
def test_open_matches_the_python_host_gpu(tmp_path, gpu_lib):
    d = str(tmp_path / "case")
    _setup(d, 0.006)
    c = cs.Case(d)
    a = sv.Solver(c.mesh, c.cfg, device=0)
    a.load_case_fields(c)
    b = sv.Solver.open(d, device=0)
    for nm in ("alpha", "U", "p_rgh", "rho", "alpha_b", "U_b", "p_rgh_b", "rho_b", "V", "Sf"):
        assert np.array_equal(a.get(nm), b.get(nm)), nm  # before the first step: copies and the same kernels
    for nm in ("owner", "neighbour", "cf", "cn", "layout"):
        assert np.array_equal(a.get_int(nm), b.get_int(nm)), nm
    a.step(4)
    b.step(4)
    assert int(b.info()["launches"]) > 0
    assert abs(a.info()["t"] - b.info()["t"]) <= 1e-12 * a.info()["t"]
    for nm, tol in (("alpha", 1e-9), ("U", 1e-7), ("p_rgh", 1e-7)):  # two handles, same configuration
        x, y = a.get(nm), b.get(nm)
        assert np.abs(x - y).max() <= tol * max(np.abs(x).max(), 1e-30), nm
    a.close()
    b.close()


@pytest.mark.gpu
def test_c_host_runs_a_case_directory_gpu(tmp_path, gpu_lib):
    """tools/tpp_foamrun.c against foamrun.run_case, both on cuda:0"""
    exe = str(tmp_path / "tpp_foamrun")
    subprocess.run(["gcc", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "tpp_foamrun.c"), "-o", exe, "-ldl"], check=True)
    a, b = str(tmp_path / "python"), str(tmp_path / "c_host")
    _setup(a, 0.006)
    _setup(b, 0.006)
    out = foamrun.run_case(a, device=0, log=None)
    r = subprocess.run([exe, gpu_lib, "-case", b], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.rstrip().splitlines()[-1] == f"End  ({out['steps']} steps)"
    assert cs.latest_time(b) == cs.latest_time(a)
    mesh = ff.read_polymesh(a)
    t = cs.latest_time(a)[1]
    for nm, tol in (("alpha.water", 1e-9), ("U", 1e-7), ("p_rgh", 1e-7), ("phi", 1e-7)):
        fa, fb = ff.read_field(os.path.join(a, t, nm)), ff.read_field(os.path.join(b, t, nm))
        n = mesh.n_internal if fa.cls.startswith("surface") else mesh.n_cells
        x, y = fa.internal_array(n), fb.internal_array(n)
        assert np.abs(x - y).max() <= tol * max(np.abs(x).max(), 1e-30), nm
    pa = open(os.path.join(a, "postProcessing", "probes", "0", "p")).read().splitlines()
    pb = open(os.path.join(b, "postProcessing", "probes", "0", "p")).read().splitlines()
    assert len(pa) == len(pb) == out["steps"] + 1 + sum(l.startswith("#") for l in pa)
    assert [l for l in pa if l.startswith("#")] == [l for l in pb if l.startswith("#")]
