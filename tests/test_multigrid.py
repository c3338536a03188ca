"""The p_rgh multigrid in isolation, through tpp_solve on a
synthetic two-phase pressure matrix (rAUf jumps 1000:1 across the free surface, Dirichlet lid):

* the result is checked against a direct sparse solve (scipy) - the solver is only allowed to
  stop on OpenFOAM's convergence contract (L1 residual / normFactor below `tolerance`);
* the persistent tail kernel (small levels in one cooperative launch) and the kernel-per-
  operation path are the same algorithm: forcing every level through either path gives the
  same iteration counts and the same solution to FP32-preconditioner round-off.
"""
import copy
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import bench
from openfoam_tpp_b200 import meshgen as mg
from openfoam_tpp_b200 import solver as sv


def _system(mesh):
    nC, nI = mesh.n_cells, mesh.n_internal
    own, nei = mesh.owner[:nI].astype(np.int64), mesh.neighbour[:nI].astype(np.int64)
    C, _ = mg.cell_geometry(mesh)
    wet = C[:, 2] <= bench.CASE["H"] / 2
    k = np.where(wet, 1e-3, 1.0)  # ~ rAU = dt/rho: air 1000 x water
    upper = 2.0 * k[own] * k[nei] / (k[own] + k[nei]) * (0.5 + ((own * 7919 + nei * 104729) % 1000) / 1000.0)
    diag = np.zeros(nC)
    np.add.at(diag, own, upper)
    np.add.at(diag, nei, upper)
    top = C[:, 2] > 0.97 * C[:, 2].max()
    diag[top] += k[top]  # the atmosphere patch (fixed value) on the lid cells
    rng = np.random.default_rng(7)
    b = rng.standard_normal(nC) * k
    A = sp.coo_matrix((np.concatenate([diag, -upper, -upper]), (np.concatenate([np.arange(nC), own, nei]), np.concatenate([np.arange(nC), nei, own]))), shape=(nC, nC)).tocsc()
    return diag, upper, b, A


def _run(lib, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        mesh = mg.cylinder_mesh(bench.CASE["H"], bench.CASE["D"], 10, 20, "flat", "tet")  # 36 000 tets
        cfg = bench.make_config(mesh)
        g = sv.Solver(mesh, cfg, device=0, lib_path=lib)
        diag, upper, b, A = _system(mesh)
        ctl = copy.copy(cfg.p_rgh_final)
        ctl.tolerance, ctl.rel_tol, ctl.max_iter = 1e-10, 0.0, 200
        x, it, r0, r = g.solve(ctl, diag, upper, b)
        lay, lev = g.amg_layout(), g.amg_levels()
        g.close()
        return x, it, r0, r, lay, lev, A, b
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _check(lib):
    x1, it1, r0, r1, lay1, lev1, A, b = _run(lib, {"TPP_TAIL_ROWS": "300000", "TPP_COARSEST": "200"})  # every coarse level in the tail kernel
    x2, it2, _, r2, lay2, lev2, _, _ = _run(lib, {"TPP_TAIL_ROWS": "200", "TPP_COARSEST": "200"})    # every coarse level kernel by kernel
    assert lay1["kernel_levels"] == 1 and lay1["tail_levels"] >= 3, lay1
    assert lay2["kernel_levels"] >= 3 and lay2["tail_levels"] == 1, lay2
    assert lev1 == lev2
    assert r1 < 1e-10 and r2 < 1e-10 and 0 < it1 < 60, (it1, r1, it2, r2)
    assert abs(it1 - it2) <= 1, (it1, it2)
    xs = spla.spsolve(A, b)
    scale = np.abs(xs).max()
    assert np.abs(x1 - xs).max() <= 1e-6 * scale, np.abs(x1 - xs).max() / scale
    assert np.abs(x2 - xs).max() <= 1e-6 * scale
    assert np.abs(x1 - x2).max() <= 1e-6 * scale


def test_tail_kernel_equals_kernel_per_operation_emu(emu_lib):
    _check(emu_lib)


@pytest.mark.gpu
def test_tail_kernel_equals_kernel_per_operation_gpu(gpu_lib):
    _check(None)


def test_gamg_as_a_stationary_iteration_emu(emu_lib):
    """`solver GAMG` (fvSolution:42-48) taken literally - V-cycles as a stationary iteration, no Krylov
    acceleration (TPP_GAMG_STATIONARY=1) - converges to the same solution under the same residual
    criterion; by default the entry runs as PCG preconditioned by the same V-cycle (fewer cycles)."""
    import copy as _copy

    its = {}
    for flag in ("0", "1"):
        os.environ["TPP_GAMG_STATIONARY"] = flag
        try:
            mesh = mg.cylinder_mesh(bench.CASE["H"], bench.CASE["D"], 10, 20, "flat", "tet")
            cfg = bench.make_config(mesh)
            g = sv.Solver(mesh, cfg, device=0, lib_path=emu_lib)
            diag, upper, b, A = _system(mesh)
            ctl = _copy.copy(cfg.p_rgh)  # type GAMG
            assert ctl.type == 1
            ctl.tolerance, ctl.rel_tol, ctl.max_iter = 1e-10, 0.0, 400
            x, it, r0, r = g.solve(ctl, diag, upper, b)
            g.close()
        finally:
            os.environ.pop("TPP_GAMG_STATIONARY", None)
        xs = spla.spsolve(A, b)
        assert r < 1e-10 and np.abs(x - xs).max() <= 1e-6 * np.abs(xs).max(), (flag, it, r)
        its[flag] = it
    assert its["0"] <= its["1"] < 400, its
