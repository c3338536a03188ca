"""`startFrom latestTime` (system/controlDict:19; main.py:336-348 picks `make resume` when a case
has progress): a run stopped at a write time and resumed from its time directory must continue
exactly like the uninterrupted run - the written state (alpha.water, U, p_rgh, phi, Uf in binary,
deltaT in uniform/time) is everything the next step needs."""
import os

import numpy as np

from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import foamfile as ff
from openfoam_tpp_b200 import foamrun


def _setup(d):
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=5, n_layers=5, write_interval=0.003)


def _steps_to(case_dir, lib, n_writes):
    """run until n_writes time directories have been written by this call"""
    c = cs.Case(case_dir)
    target = c.start_value + n_writes * c.cfg.write_interval
    p = os.path.join(case_dir, "system", "controlDict")
    s = open(p).read()
    import re

    s = re.sub(r"endTime\s+[^;]+;", f"endTime         {target!r};", s, count=1)
    open(p, "w").write(s)
    return foamrun.run_case(case_dir, lib_path=lib, log=None)


def test_resume_continues_the_uninterrupted_run(tmp_path, emu_lib):
    a, b = str(tmp_path / "straight"), str(tmp_path / "resumed")
    _setup(a)
    _setup(b)
    out_a = _steps_to(a, emu_lib, 2)
    out_b1 = _steps_to(b, emu_lib, 1)
    assert cs.latest_time(b)[1] != "0"
    out_b2 = _steps_to(b, emu_lib, 1)              # picks up the latest time directory
    assert out_b1["steps"] + out_b2["steps"] == out_a["steps"]
    ta, tb = cs.latest_time(a), cs.latest_time(b)
    assert ta == tb
    mesh = ff.read_polymesh(a)
    for nm in ("alpha.water", "U", "p_rgh", "phi", "Uf", "p", "rho"):
        fa, fb = ff.read_field(os.path.join(a, ta[1], nm)), ff.read_field(os.path.join(b, tb[1], nm))
        n = mesh.n_internal if fa.cls.startswith("surface") else mesh.n_cells
        x, y = fa.internal_array(n), fb.internal_array(n)
        err = np.abs(x - y).max() / max(np.abs(x).max(), 1e-300)
        assert err <= 1e-9, (nm, err)
    da = ff.read_dict(os.path.join(a, ta[1], "uniform", "time"))
    db = ff.read_dict(os.path.join(b, tb[1], "uniform", "time"))
    assert abs(ff.to_float(da["deltaT"]) - ff.to_float(db["deltaT"])) <= 1e-12 * ff.to_float(da["deltaT"])
    # probes: the resumed run appends a second block under its own start time, as OpenFOAM does
    assert os.path.isdir(os.path.join(b, "postProcessing", "probes"))


WORKER = """
import os, sys
sys.path.insert(0, {root!r})
from openfoam_tpp_b200 import foamrun
out = foamrun.run_case({case!r}, lib_path={lib!r}, parallel=True, log=None)
sys.stdout.write('RANK%sOK %d\\\\n' % (os.environ['RANK'], out['steps'])); sys.stdout.flush()
import torch.distributed as dist
dist.destroy_process_group()
"""


def test_parallel_resume_matches_the_serial_run(tmp_path, emu_lib):
    """`make resume N_CPUS=2` (Makefile:88-99): the processor directories of a stopped decomposed
    run are picked up again; after reconstructPar the result is the uninterrupted serial run."""
    import re
    import subprocess
    import sys
    import textwrap

    from openfoam_tpp_b200 import decompose as dc

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a, b = str(tmp_path / "serial"), str(tmp_path / "par")
    for d in (a, b):
        _setup(d)
        p = os.path.join(d, "system", "fvSolution")  # tight solves: serial and decomposed runs then agree closely
        s = open(p).read().replace("tolerance       1e-08;", "tolerance       1e-13;").replace("tolerance       2e-09;", "tolerance       1e-13;").replace("relTol          0.01;", "relTol          0;").replace("maxIter         20;", "maxIter         400;")
        open(p, "w").write(s)
    _steps_to(a, emu_lib, 2)
    with open(os.path.join(b, "system", "decomposeParDict"), "w") as f:
        f.write(ff._hdr("dictionary", "decomposeParDict", "system") + "numberOfSubdomains 2;\nmethod simple;\nsimpleCoeffs { n (2 1 1); delta 0.001; }\n" + ff.END)
    dc.decompose_par(b)
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(WORKER.format(root=root, case=b, lib=emu_lib)))
    wi = cs.Case(b).cfg.write_interval
    for k in (1, 2):
        p = os.path.join(b, "system", "controlDict")
        s = re.sub(r"endTime\s+[^;]+;", f"endTime         {k * wi!r};", open(p).read(), count=1)
        open(p, "w").write(s)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29645", str(script)],
                           capture_output=True, text=True, timeout=900)
        o = r.stdout + r.stderr
        assert r.returncode == 0 and "RANK0OK" in o and "RANK1OK" in o, o[-3000:]
    tn = cs.latest_time(a)[1]
    dc.reconstruct_par(b, [tn])
    mesh = ff.read_polymesh(a)
    for nm, tol in (("alpha.water", 1e-8), ("U", 1e-6), ("p_rgh", 1e-6), ("phi", 1e-6)):
        fa, fb = ff.read_field(os.path.join(a, tn, nm)), ff.read_field(os.path.join(b, tn, nm))
        n = mesh.n_internal if fa.cls.startswith("surface") else mesh.n_cells
        x, y = fa.internal_array(n), fb.internal_array(n)
        err = np.abs(x - y).max() / max(np.abs(x).max(), 1e-300)
        assert err <= tol, (nm, err)
