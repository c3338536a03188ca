"""The oracle and the host I/O pinned against the reference's committed OpenFOAM-13 artefacts
(G1 alpha files, G2 probes/time-step sequences, G3 interface statistics, G5
analytic potential flow) and against outputs of the reference's own Python generators.
Fixtures: tests/golden/ (made by tests/golden/make_golden.py inside the build container).

What these can and cannot pin: none of the artefacts includes the mesh, so there is no
bit-level known-answer test for the solver (PARITY UNPINNED); they pin file
formats, the time-step controller, and run statistics.
"""
import json
import os

import numpy as np
import pytest

from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import foamfile as ff
from openfoam_tpp_b200 import motion

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def test_g1_binary_field_roundtrip(tmp_path):
    """OpenFOAM-13's own binary volScalarField: read, check, rewrite, reread bit-exact."""
    src = os.path.join(HERE, "golden", "alpha.water.G1")
    f = ff.read_field(src)
    ref = G["G1_alpha"]["case_H0.004_D0.0221_flat_R0.005_f2.0"]
    assert f.cls == "volScalarField" and f.internal.size == ref["n"] == 7766
    assert float(f.internal.sum()) == ref["sum"] == 3886.0
    assert sorted(set(np.unique(f.internal).tolist())) == [0.0, 1.0]
    assert {k: v["type"] for k, v in f.boundary.items()} == ref["boundary"]
    out = tmp_path / "0" / "alpha.water"
    ff.write_field(str(out), f, binary=True, location="0")
    g = ff.read_field(str(out))
    assert np.array_equal(g.internal, f.internal)
    assert {k: v["type"] for k, v in g.boundary.items()} == ref["boundary"]
    # same payload bytes as OpenFOAM wrote
    raw_src, raw_out = open(src, "rb").read(), open(out, "rb").read()
    i, j = raw_src.index(b"\n7766\n(") + 7, raw_out.index(b"\n7766\n(") + 7
    assert raw_src[i : i + 7766 * 8] == raw_out[j : j + 7766 * 8]


def test_g1_cell_counts_recorded():
    assert [v["n"] for v in G["G1_alpha"].values()] == [7766, 18964, 41895]
    assert [v["sum"] for v in G["G1_alpha"].values()] == [3886.0, 9441.0, 20996.0]


def _oracle_case(tmp_path, **kw):
    import oracle

    d = str(tmp_path / "c")
    cs.setup_case(d, **kw)
    c = cs.Case(d)
    o = oracle.Oracle(c.mesh, c.cfg)
    o.load_case_fields(c)
    return c, o


def test_g2_first_time_steps_match_openfoam(tmp_path):
    """deltaT 0.001 -> x1.2 -> equalised to the next write time: every committed OpenFOAM run
    starts 0, 0.00119048 (= 0.05/42), independent of the mesh; then Co collapses deltaT to a few
    1e-5 s (first-step Co >> 1 from the impulsive start) and it regrows."""
    c, o = _oracle_case(tmp_path, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=10.0, n_rings=8, n_layers=4)
    ts = [0.0]
    for _ in range(12):
        o.step(1)
        ts.append(o.info()["t"])
    for case, g in G["G2_probes"].items():
        ref = g["first_times"]
        assert ref[0] == 0.0 and ref[1] == 0.00119048
        assert float(f"{ts[1]:.6g}") == ref[1], "first step must be 0.05/42 exactly as OpenFOAM prints it"
        # reference signature: second step collapses below 1e-4 s, then grows by <= 1.2x
        d_ref = np.diff(ref[:12])
        # (times are printed with 6 significant digits: allow their rounding in the ratio)
        assert d_ref[1] < 1e-4 and np.all(d_ref[2:] / d_ref[1:-1] <= 1.2 + 1e-2)
    d = np.diff(ts)
    assert abs(d[0] - 0.05 / 42) < 1e-15
    assert d[1] < 1.5e-4, "oracle must show the same start-up collapse of deltaT"
    assert np.all(d[2:] / d[1:-1] <= 1.2 + 1e-9), "growth limited to 1.2x per step"
    # Courant control: after the collapse the step follows maxCo 0.5
    assert 0.2 < o.info()["Co"] <= 0.6


def test_g2_write_times_are_hit_exactly(tmp_path):
    """adjustableRunTime: OpenFOAM lands on every multiple of writeInterval (golden: first 50
    write times hit to 4e-16)."""
    c, o = _oracle_case(tmp_path, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=5, n_layers=3, write_interval=0.002, end_time=0.0061)
    hits = []
    for _ in range(3):
        assert o.run_to_write(5000) == 1
        hits.append(o.info()["t"])
    assert np.allclose(hits, [0.002, 0.004, 0.006], rtol=0, atol=1e-15)
    assert o.run_to_write(5000) == 0  # endTime reached


def test_g2_probes_file_layout(tmp_path):
    """postProcessing/probes/0/p byte layout (header, 14-char columns, -vGreat sentinel)."""
    from openfoam_tpp_b200 import foamrun as fr

    d = str(tmp_path / "c")
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=5, n_layers=3)
    c = cs.Case(d)
    w = fr.ProbesWriter(c, "0", "p")
    ref = G["G2_probes"]["case_H0.004_D0.0221_flat_R0.005_f2.0"]
    rows = [[t, -1.79769e307, -1.79769e307] for t in ref["first_times"][:5]]
    w.rows(rows)
    w.close()
    got = open(os.path.join(d, "postProcessing", "probes", "0", "p")).readlines()
    assert got == ref["head"], f"probes layout differs:\n{got}\n{ref['head']}"


def test_motion_table_matches_reference_generator(tmp_path):
    """orbital_table restates generate_motion.py; the fixture is that script's own output."""
    rows = motion.orbital_table(0.005, 2.0, 0.05, 0.001, 0.02)
    p = tmp_path / "6DoF.dat"
    motion.write_table(str(p), rows)
    assert open(p).read() == G["motion_table_text"]
    back = motion.read_table(str(p))
    assert back.shape == (51, 7) and np.array_equal(back, rows)
    # OpenFOAM Table semantics: linear between rows, clamped outside
    mid = motion.interpolate(rows, 0.0015)
    assert np.allclose(mid, 0.5 * (rows[1, 1:] + rows[2, 1:]))
    assert np.array_equal(motion.interpolate(rows, 99.0), rows[-1, 1:])


def test_g5_potential_flow_constants():
    """First natural frequency of the D = 0.2 m tank, restated: omega^2 = g k tanh(k d),
    k = 1.8412/R  (utils/potential_flow.py:21-68); fixture computed by importing that file."""
    R, d = G["G5_potential"]["R"], G["G5_potential"]["d"]
    k = 1.8412 / R
    w = np.sqrt(9.81 * k * np.tanh(k * d))
    assert abs(w - G["G5_potential"]["omega_1n"][0]) < 1e-9
    assert abs(G["G5_potential"]["A_PT"] - 3.14693958e-02) < 1e-9


def test_g3_reference_run_statistics():
    """The committed OpenFOAM run (41 895 cells, 20 s): mean interface height stays at the fill
    level 0.104 m to within 0.2 mm +- 0.2 mm - the volume-conservation proxy the GPU run is
    compared with in the integration run (profiles/validation_r1.md)."""
    m = np.array(G["G3_interface"]["mean_z"])
    assert G["G3_interface"]["n"] == 401
    assert abs(m.mean() - 0.10418) < 2e-5 and m.std() < 3e-4


def test_committed_validation_series_against_g3():
    """profiles/validation_r1.md in numbers: the GPU run of the reference's D = 0.2 m case measured
    with the reference's own iso-surface metric reproduces OpenFOAM's t = 0 spread and ramp-up
    (golden G3), and its m = 1 amplitude settles near the analytic potential-flow value (G5)."""
    import csv

    g3 = G["G3_interface"]
    path = os.path.join(os.path.dirname(HERE), "profiles", "validation_r1_gpu_10x22_20s_v2.csv")
    rows = list(csv.DictReader(open(path)))
    col = lambda k: np.array([float(r[k]) for r in rows])
    t = col("time")
    assert len(rows) == g3["n"] == 401 and np.allclose(t, g3["t"], atol=1e-9)
    ofmax, ofmin = np.array(g3["max_z"]), np.array(g3["min_z"])
    mx, mn = col("iso_max_z"), col("iso_min_z")
    # flat surface at t = 0: the spread is the metric's own (one cell), the same on both meshes
    assert abs(mx[0] - ofmax[0]) < 1e-3 and abs(mn[0] - ofmin[0]) < 1e-3
    # ramp-up (0-2 s) and first beat maximum (2-4 s): envelopes within 15 %
    for lo, hi in ((0, 40), (40, 80)):
        a, b = (mx[lo:hi] - 0.104).max(), (ofmax[lo:hi] - 0.104).max()
        assert abs(a - b) < 0.15 * b, (lo, a, b)
    # long time: the m = 1 wall amplitude settles within 20 % of linear theory (31.5 mm)
    A = col("A_m1")[-80:]
    apt = 0.03146939582401524
    assert abs(A.mean() - apt) < 0.2 * apt and A.std() < 0.1 * apt
