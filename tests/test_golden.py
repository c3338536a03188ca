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


def _g4():
    a = np.genfromtxt(os.path.join(HERE, "golden", "g4_m1_series.csv"), delimiter=",", names=True)
    return a


def test_g4_fixture_is_a_two_mode_beat():
    """Golden G4 (the reference's 401 committed OpenFOAM-13 iso-surfaces, reduced by make_golden.py):
    after the ramp the m = 1 wall mode is the sum of the forced response at 1.88 Hz and a slowly
    decaying free mode - fitted to 0.8 mm rms over 18 s.  These four numbers are the reference's
    answer for the case; they also pin `interface.beat_fit`."""
    from openfoam_tpp_b200 import interface

    g = _g4()
    assert len(g) == 401 and np.allclose(g["time"], np.arange(401) * 0.05, atol=1e-9)
    # the same files give the reference's own summary statistics (golden G3)
    g3 = G["G3_interface"]
    assert np.allclose(g["max_z"], g3["max_z"], atol=2e-7) and np.allclose(g["min_z"], g3["min_z"], atol=2e-7)
    f = interface.beat_fit(g["time"], g["A_m1"], g["phase_m1"], 1.88, 2.0)
    assert abs(f["f0"] - 2.2085) < 0.003, f            # 5.5 % above the analytic 2.093 Hz (G5): the 9 mm tet mesh
    assert abs(f["gamma"] - 0.059) < 0.006, f
    assert abs(f["A_forced"] - 0.0173) < 0.0005 and abs(f["A_free"] - 0.0163) < 0.0008, f
    assert f["rms"] < 1.0e-3, f
    assert abs(g["A_m1"][200:].mean() - 0.01794) < 2e-4   # mean 17.9 mm over 10-20 s


def test_committed_run_on_an_unstructured_mesh_against_g4():
    """The reference case run by the CUDA solver for the full 20 s on an unstructured Delaunay tet
    mesh of the reference's size (41 535 tets at lc = 9 mm; gmsh's own: 41 895), committed as
    profiles/r2_physics/gpu_unstructured_lc9_seed0_20s.csv, against golden G4 with the same
    reduction.  The meshes are different realisations (gmsh is not available), so the comparison is
    of the quantities a mesh realisation leaves alone; BASELINE.json's 1 % / 1 % needs the identical
    mesh and is NOT claimed.  Stated tolerances = what this run meets with some margin:
      natural frequency 1.5 %, free-mode damping rate 25 %, forced amplitude 10 %, free amplitude 30 %,
      mean m = 1 amplitude over 10-20 s 15 %; over the first 10 forcing periods amplitude rms 15 % of
      its maximum and phase rms 0.05 period; step count within 10 % of OpenFOAM's 60 794."""
    from openfoam_tpp_b200 import interface

    g = _g4()
    r = np.genfromtxt(os.path.join(os.path.dirname(HERE), "profiles", "r2_physics", "gpu_unstructured_lc9_seed0_20s.csv"), delimiter=",", names=True)
    assert len(r) == 401 and np.allclose(r["time"], g["time"], atol=1e-9)
    fg = interface.beat_fit(g["time"], g["A_m1"], g["phase_m1"], 1.88, 2.0)
    fr = interface.beat_fit(r["time"], r["iso_A_m1"], r["iso_phase_m1"], 1.88, 2.0)
    assert abs(fr["f0"] / fg["f0"] - 1) < 0.015, (fr, fg)
    assert abs(fr["gamma"] / fg["gamma"] - 1) < 0.25, (fr, fg)
    assert abs(fr["A_forced"] / fg["A_forced"] - 1) < 0.10, (fr, fg)
    assert abs(fr["A_free"] / fg["A_free"] - 1) < 0.30, (fr, fg)
    assert fr["rms"] < 1.5e-3   # a linear two-mode beat to 20 s, like the reference's (no amplitude run-away)
    assert abs(r["iso_A_m1"][200:].mean() / g["A_m1"][200:].mean() - 1) < 0.15
    m = g["time"] <= 10 / 1.88  # the first 10 forcing periods
    amax = g["A_m1"][m].max()
    assert np.sqrt(np.mean((r["iso_A_m1"][m] - g["A_m1"][m]) ** 2)) < 0.15 * amax
    w = m & (g["A_m1"] > 0.3 * amax)  # the phase of a vanishing amplitude is noise
    dph = np.angle(np.exp(1j * (r["iso_phase_m1"][w] - g["phase_m1"][w])))
    assert np.sqrt(np.mean(dph**2)) / (2 * np.pi) < 0.05
    # the ramp-up is mesh-independent: amplitude at t = 1 s within 5 %
    assert abs(r["iso_A_m1"][20] / g["A_m1"][20] - 1) < 0.05
    assert abs(r["step"][-1] / 60794 - 1) < 0.10  # adaptive time stepping: G2 row count of the same case
    # volume: the iso-surface mean height stays at the fill level like the reference's (G3: 0.10418)
    assert abs(r["iso_mean_z"].mean() - g["mean_z"].mean()) < 1.0e-3


def test_mesh_realisation_moves_the_natural_frequency():
    """The experiment behind the tolerances above (profiles/r2_physics/README.md): the CPU oracle on the
    repo's structured tet mesh (39 600 cells) and on the unstructured Delaunay mesh, same case, 6.5 s.
    The mesh alone moves the fitted natural frequency from 2.14 Hz to 2.19 Hz (OpenFOAM on its gmsh
    mesh: 2.21 Hz; analytic: 2.093 Hz) and halves the damping rate to the reference's; the
    unverifiable restatement choices (clipping the compressed face value, own-cell MULES extrema,
    no ddtCorr) move it by less than 0.2 %."""
    from openfoam_tpp_b200 import interface

    d = os.path.join(os.path.dirname(HERE), "profiles", "r2_physics")
    fit = lambda nm: interface.beat_fit(*(lambda a: (a["time"], a["iso_A_m1"], a["iso_phase_m1"]))(np.genfromtxt(os.path.join(d, nm), delimiter=",", names=True)), 1.88, 2.0)
    g = _g4()
    fg = interface.beat_fit(g["time"], g["A_m1"], g["phase_m1"], 1.88, 2.0, 6.5)
    fs, fu = fit("oracle_structured_10x22_6p5s.csv"), fit("oracle_unstructured_lc9_seed0_6p5s.csv")
    assert abs(fs["f0"] - 2.142) < 0.004 and abs(fu["f0"] - 2.191) < 0.004 and abs(fg["f0"] - 2.211) < 0.004
    assert abs(fu["gamma"] / fg["gamma"] - 1) < 0.1 and fs["gamma"] > 1.8 * fg["gamma"]
    for nm in ("clip", "ownextrema", "ddtcorr0"):
        ft = fit(f"oracle_unstructured_lc9_seed0_{nm}_6p5s.csv")
        assert abs(ft["f0"] / fu["f0"] - 1) < 0.002 and abs(ft["A_forced"] / fu["A_forced"] - 1) < 0.02, (nm, ft)







@pytest.mark.gpu
def test_live_ramp_up_on_an_unstructured_mesh_against_g4(gpu_lib):
    """The CUDA solver on the unstructured Delaunay mesh (41 535 tets, the committed 20 s run's mesh)
    through the shaker's ramp (first 1.88 forcing periods, to t = 1 s): the m = 1 wall amplitude and
    phase of the reference's OpenFOAM run (golden G4) within 5 % / 0.03 period, and the run reproduces
    the committed series of the same mesh."""
    import bench
    from openfoam_tpp_b200 import case as cs
    from openfoam_tpp_b200 import foamfile as ff
    from openfoam_tpp_b200 import interface, meshgen
    from openfoam_tpp_b200 import solver as sv
    import tempfile

    C = bench.CASE
    mesh = meshgen.unstructured_cylinder_mesh(C["H"], C["D"], 0.009, seed=0, iters=40)
    assert mesh.n_cells == 41535
    with tempfile.TemporaryDirectory() as tmp:
        cs.write_template(tmp, end_time=1.0, write_interval=0.05, fill_z=C["H"] / 2)
        motion.write_table(os.path.join(tmp, "constant", "6DoF.dat"), motion.orbital_table(C["R"], C["freq"], 1.5, C["dt"], C["ramp"]))
        cfg = cs.read_config(tmp, None)
        fields = {n: ff.read_field(os.path.join(tmp, "0", n)) for n in ("U", "alpha.water", "p_rgh")}
        cs._bc_tables(cfg, mesh, fields, "0")
    cfg.start_time = 0.0
    s = sv.Solver(mesh, cfg)
    s.set("alpha", bench.initial_alpha(mesh))
    s.init_fields()
    edges = interface.mesh_edges(mesh)
    g = _g4()
    ref = np.genfromtxt(os.path.join(os.path.dirname(HERE), "profiles", "r2_physics", "gpu_unstructured_lc9_seed0_20s.csv"), delimiter=",", names=True)
    k = 0
    while s.run_to_write() == 1:
        k += 1
        if k % 5:
            continue
        iso = interface.iso_points(mesh, mesh.points, interface.cell_to_point(mesh, s.get("alpha")), 0.5, edges)
        A, ph, _, _ = interface.iso_wall_mode1(iso, (0.0, 0.0), 0.5 * C["D"])
        assert abs(A - ref["iso_A_m1"][k]) <= 0.02 * max(ref["iso_A_m1"][k], 1e-4), (k, A, ref["iso_A_m1"][k])
        if g["A_m1"][k] > 2e-3:
            assert abs(A / g["A_m1"][k] - 1) < 0.05, (k, A, g["A_m1"][k])
            assert abs(np.angle(np.exp(1j * (ph - g["phase_m1"][k])))) / (2 * np.pi) < 0.03, (k, ph, g["phase_m1"][k])
    assert k == 20 and abs(s.info()["t"] - 1.0) < 1e-9
    st = s.info()
    assert abs(st["step"] / ref["step"][20] - 1) < 0.10  # (the step size follows the fastest air cell: sensitive to the solver tolerances)


def test_cuda_and_oracle_series_agree_within_one_percent():
    """BASELINE.json's floating-point gate between the CUDA path and the CPU oracle (the OpenFOAM
    restatement), on committed series of the reference case on the same unstructured 41 535-tet mesh:
    over 18.8 forcing periods (10 s, ~32 900 adaptive steps, the reference's solver tolerances, two
    different linear solvers) the m = 1 interface amplitude agrees within 1 % of its maximum (0.4 %
    measured), its phase within 1 % of a period (0.2 %), the total water volume to 1e-6, the step
    count within 1 %.  The oracle is deterministic: its 10 s run repeats its earlier 6.5 s run exactly."""
    d = os.path.join(os.path.dirname(HERE), "profiles", "r2_physics")
    g = np.genfromtxt(os.path.join(d, "gpu_unstructured_lc9_seed0_20s.csv"), delimiter=",", names=True)
    o = np.genfromtxt(os.path.join(d, "oracle_unstructured_lc9_seed0_10s.csv"), delimiter=",", names=True)
    o65 = np.genfromtxt(os.path.join(d, "oracle_unstructured_lc9_seed0_6p5s.csv"), delimiter=",", names=True)
    n = len(o)
    assert n == 201 and o["time"][-1] * 1.88 > 18 and np.allclose(g["time"][:n], o["time"], atol=1e-9)
    for k in ("iso_A_m1", "iso_phase_m1", "alpha_volume", "step"):
        assert np.array_equal(o[k][:len(o65)], o65[k])
    A = o["iso_A_m1"]
    assert np.abs(g["iso_A_m1"][:n] - A).max() < 0.01 * A.max()
    w = A > 0.1 * A.max()
    dph = np.angle(np.exp(1j * (g["iso_phase_m1"][:n] - o["iso_phase_m1"])))
    assert np.abs(dph[w]).max() / (2 * np.pi) < 0.01
    assert np.abs(g["alpha_volume"][:n] / o["alpha_volume"] - 1).max() < 1e-6  # (what leaves through the open top differs in the last digits of the air-side traces)
    assert abs(g["step"][n - 1] / o["step"][-1] - 1) < 0.01
