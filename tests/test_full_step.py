"""Whole time steps through the C-ABI against the oracle, plus the size-independent
properties the domain offers (boundedness, volume conservation up to boundary fluxes).

Tolerances (FP64):
  * both sides solve p_rgh to 1e-13 (normalised L1 residual): the reference's own loose
    first-corrector tolerance (fvSolution:46 relTol 0.01) makes two valid solvers differ by
    ~1e-2 in U, so algorithmic parity is shown with tight solves: time-step sequence equal to
    1e-12 relative, alpha to 1e-9, U and p_rgh to 1e-8 / 1e-7 of their scale;
  * with the reference's tolerances the comparison is on integrated quantities:
    total water volume (exactly conserved up to the boundary flux) and bounded alpha.
"""
import numpy as np
import pytest

from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import solver as sv


def _tight(cfg):
    for s in (cfg.p_rgh, cfg.p_rgh_final):
        s.tolerance, s.rel_tol, s.max_iter = 1e-13, 0.0, 500


def _full_steps(case_dir, lib, n_steps, geo="flat", cell="tet", n_rings=8):
    import oracle

    cs.setup_case(case_dir, H=0.004, D=0.0221, geo=geo, R=0.005, freq=2.0, duration=1.0, n_rings=n_rings, n_layers=4, cell=cell)
    c = cs.Case(case_dir)
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
        assert abs(gi["dt"] - oi["dt"]) <= 1e-9 * oi["dt"]
        for nm, tol in (("alpha", 1e-9), ("U", 1e-8), ("p_rgh", 1e-7), ("phi", 1e-8), ("p", 1e-7)):
            a, b = g.get(nm), o.get(nm)
            err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
            assert err <= tol, f"step {i}: {nm} differs by {err:.2e} of its scale (> {tol})"
    g.close()
    o.close()


def test_full_steps_emu(tmp_path, emu_lib):
    _full_steps(str(tmp_path / "c"), emu_lib, 6)


def test_full_steps_cap_prism_emu(tmp_path, emu_lib):
    _full_steps(str(tmp_path / "c"), emu_lib, 4, geo="cap", cell="prism")


@pytest.mark.gpu
def test_full_steps_gpu(tmp_path, gpu_lib):
    _full_steps(str(tmp_path / "c"), gpu_lib, 6)


@pytest.mark.gpu
def test_full_steps_bench_layout_gpu(gpu_lib):
    """The hierarchy layout the bench runs in - mesh level (two-rows-per-thread ELL kernels), an
    ELL + overflow level of > 60 k rows smoothed kernel by kernel, the persistent tail kernel below -
    on a 0.41 M-cell mesh of the bench's own tank, against the oracle (OpenFOAM-style sequential
    GAMG), both solving tightly: time-step sequence and fields as in the small-mesh test."""
    import bench
    import oracle

    mesh, nr, nl = bench.mesh_for(4.1e5)
    cfg = bench.make_config(mesh)
    _tight(cfg)
    g = sv.Solver(mesh, cfg)
    o = oracle.Oracle(mesh, cfg)
    a0 = bench.initial_alpha(mesh)
    g.set("alpha", a0); g.init_fields()
    o.set("alpha", a0); o.stage("alphaBCs"); o.stage("mixture")
    for i in range(2):
        g.step(1); o.step(1)
        gi, oi = g.info(), o.info()
        assert abs(gi["t"] - oi["t"]) <= 1e-12 * oi["t"] and abs(gi["dt"] - oi["dt"]) <= 1e-9 * oi["dt"]
        for nm, tol in (("alpha", 1e-9), ("U", 1e-7), ("p_rgh", 1e-7), ("phi", 1e-7)):
            a, b = g.get(nm), o.get(nm)
            err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
            assert err <= tol, f"step {i}: {nm} differs by {err:.2e} of its scale (> {tol})"
    lay = g.amg_layout()
    assert lay["kernel_levels"] >= 2 and lay["tail_levels"] >= 2, lay
    g.close(); o.close()


@pytest.mark.gpu
def test_full_steps_cap_prism_gpu(tmp_path, gpu_lib):
    _full_steps(str(tmp_path / "c"), gpu_lib, 4, geo="cap", cell="prism", n_rings=10)


def _tutorial(case_dir, lib, n_steps):
    """sloshingTank3D6DoF (BASELINE.json config 2): closed hex tank, single `wall` patch, table
    with rotations up to 30 degrees -> rotating geometry, swept-volume mesh flux, wall velocity
    from the rigid transform, and the p_rgh reference cell (pRefPoint/pRefValue, fvSolution:85-86)."""
    import oracle

    cs.setup_tutorial_case(case_dir, nx=6, ny=10, nz=9)
    c = cs.Case(case_dir)
    _tight(c.cfg)
    g = sv.Solver(c.mesh, c.cfg, lib_path=lib)
    g.load_case_fields(c)
    o = oracle.Oracle(c.mesh, c.cfg)
    o.load_case_fields(c)
    assert g.info()["refCell"] == o.info()["refCell"] >= 0
    for i in range(n_steps):
        g.step(1)
        o.step(1)
        assert abs(g.info()["t"] - o.info()["t"]) <= 1e-12 * o.info()["t"]
        for nm, tol in (("meshPhi", 1e-10), ("alpha", 1e-10), ("U", 1e-9), ("p_rgh", 1e-9), ("p", 1e-9), ("phi", 1e-9)):
            a, b = g.get(nm), o.get(nm)
            err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
            assert err <= tol, f"step {i}: {nm} differs by {err:.2e} of its scale (> {tol})"
        assert abs(g.get("p")[int(g.info()["refCell"])] - c.cfg.p_ref_value) < 1e-6
    # rigid motion conserves the swept volume: sum of meshPhi over each cell's faces = 0
    nI = c.mesh.n_internal
    mp = g.get("meshPhi")
    div = np.zeros(c.mesh.n_cells)
    np.add.at(div, c.mesh.owner, mp)
    np.add.at(div, c.mesh.neighbour, -mp[:nI])
    assert np.abs(div).max() <= 1e-9 * np.abs(mp).max()
    g.close()
    o.close()


def test_tutorial_tank_rotation_emu(tmp_path, emu_lib):
    _tutorial(str(tmp_path / "c"), emu_lib, 5)


@pytest.mark.gpu
def test_tutorial_tank_rotation_gpu(tmp_path, gpu_lib):
    _tutorial(str(tmp_path / "c"), gpu_lib, 5)


def _conservation(case_dir, lib, n_steps, n_rings, n_layers):
    """Reference tolerances.  Sum(alpha V) may change only through the boundary flux
    alphaPhi_b: |d(sum alpha V) + dt*sum_b alphaPhi_b| <= 1e-10 * sum(alpha V) per step
    (BASELINE.json: total alpha volume conserved to 1e-10 relative); 0 <= alpha <= 1."""
    cs.setup_case(case_dir, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=n_rings, n_layers=n_layers)
    c = cs.Case(case_dir)
    g = sv.Solver(c.mesh, c.cfg, lib_path=lib)
    g.load_case_fields(c)
    V = g.get("V")
    nI = c.mesh.n_internal
    vol0 = float((g.get("alpha") * V).sum())
    vol = vol0
    for i in range(n_steps):
        g.step(1)
        a = g.get("alpha")
        dt = g.info()["dt"]
        out = float(g.get("alphaPhi")[nI:].sum()) * dt
        new = float((a * V).sum())
        assert abs(new - vol + out) <= 1e-10 * vol0, f"step {i}: water volume not conserved ({new - vol + out:.3e})"
        assert a.min() >= -1e-9 and a.max() <= 1 + 1e-6, f"step {i}: alpha out of bounds [{a.min()}, {a.max()}]"
        vol = new
        gi = g.info()
        assert gi["r1"] < c.cfg.p_rgh_final.tolerance or gi["it1"] >= c.cfg.p_rgh_final.max_iter
    g.close()


def test_conservation_emu(tmp_path, emu_lib):
    _conservation(str(tmp_path / "c"), emu_lib, 12, 8, 4)


@pytest.mark.gpu
def test_conservation_gpu(tmp_path, gpu_lib):
    _conservation(str(tmp_path / "c"), gpu_lib, 25, 24, 12)
