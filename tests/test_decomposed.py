"""One case decomposed over ranks: z-slabs with `processor` patches, halo
exchange before every stencil kernel, all-reduced Courant numbers / Krylov dots / residuals,
a global multigrid (processor interfaces agglomerated level by level, smallest levels gathered
onto every rank) inside a global PCG.

world_size 2, gloo, host emulation of the kernels (CPU CI of the N > 1 logic): the decomposed run
must reproduce the whole-mesh single-rank run - identical time-step sequence, fields to the
tight-solve tolerance.  The `gpu2` variant runs the same check over NCCL on two B200s
(`gpurun --gpus 2 -- python -m pytest tests -m gpu2`)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = """
import os, sys, numpy as np
sys.path.insert(0, {root!r}); sys.argv = ['x']
import torch, torch.distributed as dist
import bench
from openfoam_tpp_b200 import meshgen as mg, solver as sv
LIB = {lib!r}
nccl = LIB is None
if nccl:
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
else:
    dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
NR, NL = {nr}, {nl}
C = bench.CASE
whole = mg.cylinder_mesh(C['H'], C['D'], NR, NL, 'flat', 'tet')
k0, k1 = rank * NL // world, (rank + 1) * NL // world
mesh = mg.cylinder_mesh(C['H'], C['D'], NR, NL, 'flat', 'tet', k0=k0, k1=k1, proc=(rank, rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None))
def tight(cfg):
    for s in (cfg.p_rgh, cfg.p_rgh_final): s.tolerance, s.rel_tol, s.max_iter = 1e-13, 0.0, 800
    cfg.sigma = {sigma}
    return cfg
g = sv.Solver(mesh, tight(bench.make_config(mesh)), device=int(os.environ.get('LOCAL_RANK', 0)), lib_path=LIB)
ng, patches = g.ghost_layout()
assert ng == sum(c for _, c, _ in patches) > 0 and all(p in (rank - 1, rank + 1) for _, _, p in patches)
g.comm_init_nccl() if nccl else g.comm_init_callbacks()
g.set('alpha', bench.initial_alpha(mesh)); g.init_fields()
gw = sv.Solver(whole, tight(bench.make_config(whole)), device=int(os.environ.get('LOCAL_RANK', 0)), lib_path=LIB)
gw.set('alpha', bench.initial_alpha(whole)); gw.init_fields()
cpl = whole.n_cells // NL
sl = slice(k0 * cpl, k1 * cpl)
for i in range({steps}):
    g.step(1); gw.step(1)
    gi, wi = g.info(), gw.info()
    assert gi['t'] == wi['t'] and gi['Co'] == wi['Co'] or abs(gi['Co'] - wi['Co']) <= 1e-9 * wi['Co'], (gi['t'], wi['t'], gi['Co'], wi['Co'])
    assert abs(gi['t'] - wi['t']) <= 1e-12 * wi['t']
    for nm, nc, tol in (('alpha', 1, 1e-9), ('U', 3, 1e-7), ('p_rgh', 1, 1e-7), ('rho', 1, 1e-9)):
        a, b = g.get(nm), gw.get(nm).reshape(-1, nc)[sl].reshape(-1)
        err = np.abs(a - b).max() / max(np.abs(gw.get(nm)).max(), 1e-300)
        assert err <= tol, (i, nm, err)
# face fields come back in OpenFOAM file order: the processor patch is last
phi = g.get('phi'); nI = mesh.n_internal
pp = [p for p in mesh.patches if p['type'] == 'processor'][0]
assert np.abs(phi[pp['startFace']:pp['startFace'] + pp['nFaces']]).max() > 0
print(f'RANK{{rank}} OK cells={{mesh.n_cells}} ghosts={{ng}} iters={{int(gi["it1"])}}/{{int(wi["it1"])}}', flush=True)
dist.destroy_process_group()
"""


def _run(tmp_path, lib, nr, nl, steps, port, nproc=2, env=None, sigma=0.0):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(WORKER.format(root=ROOT, lib=lib, nr=nr, nl=nl, steps=steps, sigma=sigma)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    for k in range(nproc):
        assert f"RANK{k} OK" in out, out[-4000:]


def test_two_rank_decomposition_matches_single_rank_gloo(tmp_path, emu_lib):
    _run(tmp_path, emu_lib, nr=5, nl=8, steps=4, port=29631)


def test_three_rank_decomposition_gloo(tmp_path, emu_lib):
    """a middle slab has two processor patches (two neighbours)"""
    _run(tmp_path, emu_lib, nr=4, nl=9, steps=3, port=29632, nproc=3)


def test_two_rank_decomposition_with_surface_tension_gloo(tmp_path, emu_lib):
    """sigma > 0 (an extension; the reference runs sigma 0): grad alpha and the curvature cross the processor
    patch by halo exchange before the face normal flux and the face force are formed"""
    _run(tmp_path, emu_lib, nr=5, nl=8, steps=4, port=29635, sigma=0.072)


def test_distributed_coarse_levels_gloo(tmp_path, emu_lib):
    """tiny TPP_TAIL_ROWS: several multigrid levels stay distributed, so the agglomerated processor
    interfaces and the per-level halo exchanges are exercised, then the gathered tail"""
    _run(tmp_path, emu_lib, nr=6, nl=12, steps=3, port=29634, nproc=3, env={"TPP_TAIL_ROWS": "300", "TPP_COARSEST": "100"})


@pytest.mark.gpu
@pytest.mark.gpu2
def test_two_rank_decomposition_matches_single_rank_nccl(tmp_path, gpu_lib):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    _run(tmp_path, None, nr=12, nl=24, steps=4, port=29633)
