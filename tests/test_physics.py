"""Physics pin that needs no OpenFOAM: free sloshing of the D = 0.2 m tank at rest.  The first
azimuthal mode started from a tilted free surface must oscillate with the natural frequency of
linear potential theory - the reference's own analytic module (`utils/potential_flow.py`,
committed as golden G5: omega_11 = 13.1508 rad/s for R = 0.1 m, d = 0.104 m).  On a 1.7 k-cell
mesh the discrete period is 2-6 % short of the analytic one (16 k cells: 1.5 %)."""
import json
import os

import numpy as np

import bench
from openfoam_tpp_b200 import interface
from openfoam_tpp_b200 import meshgen as mg
from openfoam_tpp_b200 import solver as sv

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))["G5_potential"]


def test_free_sloshing_period_matches_potential_flow_emu(emu_lib):
    R, d, H = GOLDEN["R"], GOLDEN["d"], 0.208
    T = 2 * np.pi / GOLDEN["omega_1n"][0]
    mesh = mg.cylinder_mesh(H, 2 * R, 4, 6, "flat", "tet")
    cfg = bench.make_config(mesh)
    cfg.n_motion, cfg.motion = 0, None                 # tank at rest
    cfg.max_delta_t = cfg.delta_t = 0.002
    C, V = mg.cell_geometry(mesh)
    a0 = 0.004
    alpha = np.clip((d + a0 * C[:, 0] / R - C[:, 2]) / (H / 6) + 0.5, 0.0, 1.0)  # surface tilted about the y axis
    g = sv.Solver(mesh, cfg, lib_path=emu_lib)
    g.set("alpha", alpha)
    g.init_fields()
    cols = interface.ColumnSampler(mesh)
    m = cols.r > 0.85 * cols.r.max()
    A = np.stack([np.ones(m.sum()), np.cos(cols.theta[m]), np.sin(cols.theta[m])], 1)

    def cos_amplitude():
        return np.linalg.lstsq(A, cols.heights(g.get("alpha"))[m], rcond=None)[0][1]

    ts, cs_ = [0.0], [cos_amplitude()]
    vol0 = float((alpha * V).sum())
    while ts[-1] < 0.62:
        g.step(1)
        ts.append(g.info()["t"])
        cs_.append(cos_amplitude())
    vol1 = float((g.get("alpha") * V).sum())
    g.close()
    ts, c = np.array(ts), np.array(cs_)
    zc = [ts[i] + (ts[i + 1] - ts[i]) * c[i] / (c[i] - c[i + 1]) for i in range(len(c) - 1) if c[i] * c[i + 1] < 0]
    assert abs(c[0] - a0) < 0.2 * a0                       # the fit sees the imposed tilt
    assert len(zc) >= 3
    assert abs(zc[0] - T / 4) < 0.06 * T                   # first quarter period
    period = zc[2] - zc[0]                                 # one full period between like crossings
    assert abs(period - T) < 0.06 * T, (period, T)
    assert c.min() < -0.6 * a0                             # it swings to the other side, weakly damped
    # the tank is open at the top (inletOutlet): on 6 layers a trace of smeared alpha reaches the
    # lid and leaves with the displaced air; the walls are tight (phi_b = 0 exactly)
    assert 0 <= vol0 - vol1 < 5e-3 * vol0


def test_hydrostatic_rest_state_stays_at_rest_emu(emu_lib):
    """A flat free surface on a cell-layer boundary of an orthogonal (hex) tank at rest is a discrete
    equilibrium of the p_rgh formulation: it must stay at rest - no spurious currents from the
    1000:1 density jump beyond what the pressure tolerance (2e-9) admits, alpha unchanged.  (On the
    tet meshes the start from p_rgh = 0 is violent in OpenFOAM too: golden G2 pins that start-up
    deltaT collapse in tests/test_golden.py.)"""
    mesh = mg.box_mesh(4, 4, 6, lo=(0, 0, 0), hi=(0.1, 0.1, 0.2), cell="hex", top_patch="atmosphere")
    cfg = bench.make_config(mesh)
    cfg.n_motion, cfg.motion = 0, None
    cfg.max_delta_t = cfg.delta_t = 0.002
    C, V = mg.cell_geometry(mesh)
    alpha = (C[:, 2] < 0.1).astype(float)
    g = sv.Solver(mesh, cfg, lib_path=emu_lib)
    g.set("alpha", alpha)
    g.init_fields()
    g.step(20)
    U, a, i = g.get("U"), g.get("alpha"), g.info()
    g.close()
    assert i["t"] > 0.03
    assert np.abs(U).max() < 1e-5, np.abs(U).max()       # m/s; the gravity-wave speed here is ~1 m/s
    assert np.abs(a - alpha).max() < 1e-6


def _linear_forced_elevation(R, a, f, d, r, f0=None, n_modes=60):
    """Linear potential theory for an orbitally shaken cylinder: amplitude of the rotating m = 1 elevation at radius r.
    The tank-frame body force a w^2 (cos wt, sin wt) is expanded in the sloshing modes J1(e_n r / R) (Dini series of r:
    r = sum_n 2 R J1(e_n r / R) / ((e_n^2 - 1) J1(e_n)), J1'(e_n) = 0), each responding with w_n^2 / (w_n^2 - w^2):
        eta(r) = F [ r + 2 R sum_n S_n J1(e_n r / R) / J1(e_n) ],  S_n = 1 / ((e_n^2 - 1)(w_n^2 / w^2 - 1)),  F = a w^2 / g,
    i.e. eta(R) = 2 R F (1/2 + sum S_n): the quasi-static tilt R F when w -> 0.  The reference's
    utils/potential_flow.py:71-116 returns 2 R F (1 + sum S_n) as `A_PT` - R F too much (2 R F for w -> 0; golden G5:
    31.47 mm where this gives 25.78 mm).  f0: replace the first natural frequency by the discrete mesh's."""
    from scipy.special import j1, jnp_zeros

    e = jnp_zeros(1, n_modes)
    w, g = 2 * np.pi * f, 9.81
    wn2 = g * e / R * np.tanh(e / R * d)
    if f0 is not None:
        wn2[0] = (2 * np.pi * f0) ** 2
    S = 1.0 / ((e**2 - 1) * (wn2 / w**2 - 1))
    F = a * w * w / g
    return F * (r + 2 * R * (S * j1(e * r / R) / j1(e)).sum()), 2 * R * F * (1 + S.sum())


def test_forced_response_follows_linear_theory_emu(emu_lib):
    """Off resonance (forcing 1.6 Hz, first mode 2.09 Hz) the D = 0.2 m tank of the reference's m0.009 case must settle
    on linear theory's rotating wave.  The forced part of the m = 1 wall series (two-mode fit, interface.beat_fit,
    residual 0.1 mm) is within 5 % of the theory evaluated with the mesh's own first natural frequency (2.3 % measured
    on 3.5 k prisms) - and 35 % below the reference's `compute_wall_amplitude`, whose constant term is R F too large
    (see _linear_forced_elevation).  With that term fixed the reference's own OpenFOAM run agrees with theory as well:
    18.4 mm predicted at its discrete f0 = 2.209 Hz, 17.3 mm in golden G4 (profiles/r2_physics/README.md)."""
    from openfoam_tpp_b200 import motion

    R, d, H, a, f = 0.1, 0.104, 0.208, 0.004, 1.6
    assert abs(_linear_forced_elevation(R, a, 1.88, d, R, n_modes=30)[1] - GOLDEN["A_PT"]) < 1e-5  # the reference's formula, restated (its root finder misplaces the zeros from the 6th on: 3e-6 m)
    assert abs(_linear_forced_elevation(R, a, 1e-3, d, R)[0] / (R * a * (2e-3 * np.pi) ** 2 / 9.81) - 1) < 1e-6  # quasi-static tilt
    mesh = mg.cylinder_mesh(H, 2 * R, 6, 12, "flat", "prism")
    cfg = bench.make_config(mesh, freq=f)
    rows = motion.orbital_table(a, f, 7.5, 0.001, 2.0)
    cfg.motion, cfg.n_motion = rows, len(rows)
    cfg.max_delta_t = 0.004
    g = sv.Solver(mesh, cfg, lib_path=emu_lib)
    g.set("alpha", bench.initial_alpha(mesh))
    g.init_fields()
    cols = interface.ColumnSampler(mesh)
    m = cols.r > 0.85 * cols.r.max()
    A = np.stack([np.ones(m.sum()), np.cos(cols.theta[m]), np.sin(cols.theta[m])], 1)
    ts, q = [], []
    while not ts or ts[-1] < 7.0:
        g.step(1)
        c = np.linalg.lstsq(A, cols.heights(g.get("alpha"))[m], rcond=None)[0]
        ts.append(g.info()["t"])
        q.append(c[1] + 1j * c[2])
    g.close()
    q = np.array(q)
    fit = interface.beat_fit(np.array(ts), np.abs(q), np.angle(q), f, t0=2.0)
    assert fit["rms"] < 2e-4 and 2.0 < fit["f0"] < 2.25, fit
    r_ring = float(cols.r[m].mean())
    eta_ring, _ = _linear_forced_elevation(R, a, f, d, r_ring, f0=fit["f0"])
    eta_wall, a_ref = _linear_forced_elevation(R, a, f, d, R, f0=fit["f0"])
    assert abs(fit["A_forced"] / eta_ring - 1) < 0.05, (fit, eta_ring)
    assert fit["A_forced"] * eta_wall / eta_ring < 0.72 * a_ref
