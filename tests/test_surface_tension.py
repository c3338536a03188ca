"""Surface tension (sigma > 0): an extension beyond the reference, which runs `sigma 0` in every
case (constant/phaseProperties:19) - so there is no reference artefact to pin it against.  What is
checked instead:
  * the CUDA path's kernels (host emulation of the same bodies here, the GPU under `-m gpu`)
    against the oracle's restatement of interfaceProperties (curvature, face force, whole steps);
  * physics the continuum-surface-force model must reproduce: the Laplace pressure jump of a
    static drop, and capillary-gravity sloshing - the first azimuthal mode of a small cylinder
    oscillates at omega^2 = (g k + sigma k^3 / rho) tanh(k d), k = 1.8412 / R (90 degree contact
    angle = the reference's zeroGradient walls, 0/alpha.water:22-25).
"""
import numpy as np
import pytest

import bench
from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import interface
from openfoam_tpp_b200 import meshgen as mg
from openfoam_tpp_b200 import solver as sv


def _tight(cfg):
    for s in (cfg.p_rgh, cfg.p_rgh_final):
        s.tolerance, s.rel_tol, s.max_iter = 1e-13, 0.0, 500


def _steps_against_oracle(case_dir, lib, n_steps, cell="tet"):
    import oracle

    cs.setup_case(case_dir, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=8, n_layers=4, cell=cell)
    c = cs.Case(case_dir)
    c.cfg.sigma = 0.072
    # tank at rest: the oracle recomputes the moved geometry every step while the library keeps the body-frame one, and
    # nHat = g / (|g| + deltaN) amplifies that round-off by 1 / deltaN wherever grad alpha vanishes; at rest both
    # sides see identical geometry and only the pressure solvers differ (moving mesh + sigma, with its looser
    # bound: tests/test_stage_parity.py)
    c.cfg.motion = None
    _tight(c.cfg)
    g = sv.Solver(c.mesh, c.cfg, lib_path=lib)
    g.load_case_fields(c)
    o = oracle.Oracle(c.mesh, c.cfg)
    o.load_case_fields(c)
    for i in range(n_steps):
        g.step(1)
        o.step(1)
        gi, oi = g.info(), o.info()
        assert abs(gi["t"] - oi["t"]) <= 1e-12 * oi["t"], f"step {i}: time diverged"
        assert np.abs(o.get("stf")).max() > 0
        # (bit-exactness of the interface kernels on a fixed state: tests/test_stage_parity.py)
        for nm, tol in (("alpha", 1e-9), ("U", 1e-8), ("p_rgh", 1e-7), ("phi", 1e-8), ("sigmaK", 1e-4), ("stf", 1e-5)):
            a, b = g.get(nm), o.get(nm)
            err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
            assert err <= tol, f"step {i}: {nm} differs by {err:.2e} of its scale (> {tol})"
    g.close()
    o.close()


def test_surface_tension_steps_emu(tmp_path, emu_lib):
    _steps_against_oracle(str(tmp_path / "c"), emu_lib, 5)


@pytest.mark.gpu
def test_surface_tension_steps_gpu(tmp_path, gpu_lib):
    _steps_against_oracle(str(tmp_path / "c"), gpu_lib, 5)


def test_case_reader_takes_sigma_from_phase_properties(tmp_path):
    """constant/phaseProperties:19 - any non-negative sigma is read; the walls stay zeroGradient (no contact angle)."""
    import os

    d = str(tmp_path / "c")
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=0.1, n_rings=3, n_layers=2)
    p = os.path.join(d, "constant", "phaseProperties")
    txt = open(p).read()
    with open(p, "w") as f:
        f.write(txt.replace("sigma           0", "sigma           0.072", 1))
    assert cs.Case(d).cfg.sigma == 0.072


def test_sigma_zero_allocates_nothing_and_is_the_reference_path(emu_lib):
    mesh = mg.cylinder_mesh(0.004, 0.0221, 3, 2, "flat", "tet")
    cfg = bench.make_config(mesh)
    assert cfg.sigma == 0.0
    g = sv.Solver(mesh, cfg, lib_path=emu_lib)
    assert g.size("stf") == -1 and g.size("sigmaK") == -1
    g.close()


def _m1_period(lib, sigma, R, H, d, a0):
    """Free sloshing of a small prism-mesh cylinder at rest from a tilted surface: the period of the first
    azimuthal mode from like zero crossings of its cosine amplitude."""
    mesh = mg.cylinder_mesh(H, 2 * R, 4, 16, "flat", "prism")
    cfg = bench.make_config(mesh)
    cfg.n_motion, cfg.motion = 0, None
    cfg.max_delta_t = cfg.delta_t = 2.5e-4
    cfg.sigma = sigma
    C, _ = mg.cell_geometry(mesh)
    alpha = np.clip((d + a0 * C[:, 0] / R - C[:, 2]) / (H / 16) + 0.5, 0.0, 1.0)
    g = sv.Solver(mesh, cfg, lib_path=lib)
    g.set("alpha", alpha)
    g.init_fields()
    cols = interface.ColumnSampler(mesh)
    m = cols.r > 0.85 * cols.r.max()
    A = np.stack([np.ones(m.sum()), np.cos(cols.theta[m]), np.sin(cols.theta[m])], 1)
    ts, c = [0.0], [np.linalg.lstsq(A, cols.heights(alpha)[m], rcond=None)[0][1]]
    while ts[-1] < 0.21:
        g.step(1)
        ts.append(g.info()["t"])
        c.append(np.linalg.lstsq(A, cols.heights(g.get("alpha"))[m], rcond=None)[0][1])
    g.close()
    zc = [ts[i] + (ts[i + 1] - ts[i]) * c[i] / (c[i] - c[i + 1]) for i in range(len(c) - 1) if c[i] * c[i + 1] < 0]
    assert len(zc) >= 3 and min(c) < -0.6 * c[0]
    return zc[2] - zc[0]


def test_capillary_gravity_sloshing_period_emu(emu_lib):
    """omega^2 = (g k + sigma k^3 / rho) tanh(k d), k = 1.8412 / R: in a D = 22.1 mm cylinder (the reference's
    small tank, main.py defaults) water's surface tension shortens the first mode's period by 9 %.  Measured on
    1.5 k prisms: each period within 2 % of the formula, their ratio within 1.5 %."""
    R, H, d, a0, sigma, rho, grav = 0.01105, 0.02, 0.01, 3e-4, 0.072, 998.2, 9.81
    k = 1.8412 / R
    T = lambda s: 2 * np.pi / np.sqrt((grav * k + s * k**3 / rho) * np.tanh(k * d))
    T0, T1 = _m1_period(emu_lib, 0.0, R, H, d, a0), _m1_period(emu_lib, sigma, R, H, d, a0)
    assert abs(T0 - T(0.0)) < 0.02 * T(0.0), (T0, T(0.0))
    assert abs(T1 - T(sigma)) < 0.02 * T(sigma), (T1, T(sigma))
    assert abs(T0 / T1 - T(0.0) / T(sigma)) < 0.015 * T(0.0) / T(sigma), (T0 / T1, T(0.0) / T(sigma))


def _laplace_jump(lib, n, width=1.5, L=0.01, R0=0.0025, sigma=0.072):
    """A static drop without gravity in a closed box of n^3 hexes: p inside minus p outside after one step."""
    mesh = mg.box_mesh(n, n, n, lo=(0, 0, 0), hi=(L, L, L), cell="hex")
    cfg = bench.make_config(mesh)
    cfg.n_motion, cfg.motion = 0, None
    cfg.g = np.zeros(3) if isinstance(cfg.g, np.ndarray) else (0.0, 0.0, 0.0)
    cfg.sigma = sigma
    cfg.p_ref_point, cfg.p_ref_value = (0.3 * L / n, 0.3 * L / n, 0.3 * L / n), 0.0
    cfg.max_delta_t = cfg.delta_t = 2e-5
    C, _ = mg.cell_geometry(mesh)
    r = np.linalg.norm(C - 0.5 * L, axis=1)
    g = sv.Solver(mesh, cfg, lib_path=lib)
    g.set("alpha", 0.5 * (1.0 - np.tanh((r - R0) / (width * L / n))))
    g.init_fields()
    g.step(1)
    p, K, a, U = g.get("p_rgh"), g.get("sigmaK") / sigma, g.get("alpha"), g.get("U")
    g.close()
    assert np.abs(U).max() < 0.01                      # parasitic currents stay small on this profile
    band = (a > 0.3) & (a < 0.7)
    return p[r < 0.4 * R0].mean() - p[r > 1.8 * R0].mean(), K[band].mean()


def test_laplace_pressure_jump_of_a_static_drop_emu(emu_lib):
    """Young-Laplace: dp = 2 sigma / R0 and K = 2 / R0 for a sphere.  With the interface resolved (a tanh profile
    1.5 cells wide) the continuum surface force converges to it: 3.4 % at 6 cells per radius, 0.3 % at 8.  (A
    one-cell-sharp alpha gives 0.82 of the jump at every resolution - the known deficiency of curvature from an
    unsmoothed volume fraction, which OpenFOAM's interfaceProperties shares.)"""
    sigma, R0 = 0.072, 0.0025
    dp24, K24 = _laplace_jump(emu_lib, 24)
    dp32, K32 = _laplace_jump(emu_lib, 32)
    lap = 2 * sigma / R0
    assert abs(dp24 - lap) < 0.05 * lap, (dp24, lap)
    assert abs(dp32 - lap) < 0.015 * lap, (dp32, lap)
    assert abs(dp32 - lap) < abs(dp24 - lap)
    assert abs(K32 - 2 / R0) < 0.04 * 2 / R0, (K32, 2 / R0)
