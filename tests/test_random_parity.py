"""Property test of the explicit kernels: on RANDOM states (hypothesis-seeded) every kernel of a
step still reproduces the CPU oracle bit for bit - random alpha in [0, 1] (sharp and smeared),
random velocity and flux fields with both signs (every limiter branch: r = +-1000 clamps, upwind /
downwind sides, MULES lambda < 1), on tets, prisms and hexes.  Host emulation of the kernel bodies
(the same source the sm_100a library is built from); the GPU counterpart of the fixed-state stage
test is tests/test_stage_parity.py::test_static_bit_exact_gpu."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import parity as P
from openfoam_tpp_b200 import case as cs
from openfoam_tpp_b200 import solver as sv

ALPHA = ["alpha", "alpha_b", "phiBD", "phiCorr", "lambda", "alphaPhiUn", "alphaPhi", "rho", "rho_b", "rhoPhi", "grad:gradAlpha"]
MOM = ["U_b", "gradU", "mLower", "mUpper", "mDiag", "mSource", "mBIC", "mBBC"]
PREP = ["rAU", "HbyA", "HbyA_b", "rAUf", "phiHbyA", "phig", "pGrad_b", "grad:gradRho"]
ASM = ["p_rgh_b", "pUpper", "pCorrFlux", "pDiag", "pSource", "grad:gradP"]
FIN = ["p_rgh_b", "phi", "U", "U_b", "Uf", "p", "p_rgh"]
_CASES = {}


def _pair(tmp, cell, geo, lib):
    key = (cell, geo)
    geo, tag = geo.split(":")
    if key not in _CASES:
        import oracle

        d = str(tmp / f"case_{cell}_{geo}_{tag}")
        if cell == "hex":
            cs.setup_tutorial_case(d, nx=4, ny=6, nz=5, end_time=1.0)  # the tutorial tank's hex block mesh
        else:
            cs.setup_case(d, H=0.004, D=0.0221, geo=geo, R=0.005, freq=2.0, duration=1.0, n_rings=4, n_layers=3, cell=cell)
        c = cs.Case(d)
        c.cfg.motion = None
        o = oracle.Oracle(c.mesh, c.cfg)
        o.load_case_fields(c)
        g = sv.Solver(c.mesh, c.cfg, lib_path=lib)
        g.load_case_fields(c)
        for st_ in ("courant", "adjustDeltaT", "advanceTime", "moveMesh"):  # same time / deltaT on both sides
            o.stage(st_)
            g.stage(st_)
        P.sync_geometry(g, o, c.mesh, c.cfg)
        _CASES[key] = (c, g, o)
    return _CASES[key]


def _exact(g, o, names, what):
    for nm in names:
        gn, on = nm.split(":") if ":" in nm else (nm, nm)
        a, b = g.get(gn), o.get(on)
        n = min(a.size, b.size)
        assert np.array_equal(a[:n], b[:n]), f"{what}: {gn} not bit-exact (max abs diff {np.abs(a[:n] - b[:n]).max():.3e})"


@pytest.mark.parametrize("cell,geo", [("tet", "flat"), ("prism", "cap"), ("hex", "tank")])
@settings(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2**31 - 1), sharp=st.booleans(), umag=st.sampled_from([1e-3, 0.05, 1.0]))
def test_random_state_step_is_bit_exact_emu(tmp_path_factory, emu_lib, cell, geo, seed, sharp, umag):
    _random_state_step(tmp_path_factory.getbasetemp(), emu_lib, "emu", cell, geo, seed, sharp, umag)


@pytest.mark.gpu
@pytest.mark.parametrize("cell,geo", [("tet", "flat"), ("prism", "cap"), ("hex", "tank")])
def test_random_state_step_is_bit_exact_gpu(tmp_path_factory, gpu_lib, cell, geo):
    """the same property on the sm_100a kernels themselves (fixed seeds: sharp and smeared alpha,
    slow and fast flow), hex cells included"""
    for seed, sharp, umag in ((11, True, 0.05), (12, False, 1.0), (13, False, 1e-3), (14, True, 1.0)):
        _random_state_step(tmp_path_factory.getbasetemp(), None, "gpu", cell, geo, seed, sharp, umag)


def _random_state_step(base, lib, tag, cell, geo, seed, sharp, umag):
    c, g, o = _pair(base, cell, geo + ":" + tag, lib)
    rng = np.random.default_rng(seed)
    nC, nF, nI = c.mesh.n_cells, c.mesh.n_faces, c.mesh.n_internal
    a = rng.random(nC)
    if sharp:
        a = (a > 0.5).astype(float)
    U = umag * rng.standard_normal((nC, 3))
    Sf = o.get("Sf").reshape(-1, 3)
    Uf = umag * rng.standard_normal((nF, 3))
    phi = (Uf * Sf).sum(axis=1)
    phi[nI:] *= rng.random(nF - nI) > 0.3  # some boundary faces closed, in- and outflow on the others
    state = {"alpha": a, "U": U.reshape(-1), "U0": U.reshape(-1) * 0.9, "p_rgh": 10.0 * rng.standard_normal(nC), "phi": phi, "Uf": Uf.reshape(-1), "Uf0": Uf.reshape(-1) * 0.8}
    for nm, v in state.items():
        o.set(nm, v)
    for st_ in ("alphaBCs", "mixture", "UBCs"):
        o.stage(st_)
    P.sync_state(g, o)
    o.stage("alphaPredictor"); g.stage("alphaPredictor")
    _exact(g, o, ALPHA, "alphaPredictor")
    P.sync_state(g, o)
    o.stage("momentum"); g.stage("momentum")
    _exact(g, o, MOM, "momentum")
    for corr in (0, 1):
        P.sync_state(g, o)
        o.stage("pcPrepare"); g.stage("pcPrepare")
        _exact(g, o, PREP, f"pcPrepare {corr}")
        o.stage("pcAssemble"); g.stage("pcAssemble")
        _exact(g, o, ASM, f"pcAssemble {corr}")
        p = o.get("p_rgh") + rng.standard_normal(nC)  # any pressure: the explicit finish must agree on it
        o.set("p_rgh", p); g.set("p_rgh", p)
        for st_ in ("pcFinish", "pcEnd"):
            o.stage(st_); g.stage(st_)
        _exact(g, o, FIN, f"pcFinish {corr}")
