"""Iteration counts of the p_rgh solves on the host emulation of the kernels (tests/_emu), whole
mesh vs z-slab decomposition over gloo ranks.  Tuning aid for the multigrid knobs (TPP_* env):

  python tools/amg_experiment.py --cells 2e5 --steps 3
  python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29650 tools/amg_experiment.py --cells 2e5
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from openfoam_tpp_b200 import meshgen as mg  # noqa: E402
from openfoam_tpp_b200 import solver as sv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=float, default=2e5, help="cells in total")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--lib", default=os.path.join(ROOT, "tests", "_emu", "libtppvof_emu.so"))
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
C = bench.CASE
ratio = C["H"] / (C["D"] / 2)
nr = max(4, int(round((args.cells / (18.0 * ratio)) ** (1.0 / 3.0))))
nl = max(world, int(round(nr * ratio / world)) * world)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("gloo")
    k0, k1 = rank * nl // world, (rank + 1) * nl // world
    mesh = mg.cylinder_mesh(C["H"], C["D"], nr, nl, "flat", "tet", k0=k0, k1=k1, proc=(rank, rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None))
else:
    mesh = mg.cylinder_mesh(C["H"], C["D"], nr, nl, "flat", "tet")
g = sv.Solver(mesh, bench.make_config(mesh), device=0, lib_path=None if args.lib == "gpu" else args.lib)
if world > 1:
    g.comm_init_callbacks()
g.set("alpha", bench.initial_alpha(mesh))
g.init_fields()
t0 = time.time()
its = []
for i in range(args.steps):
    g.step(1)
    gi = g.info()
    its.append((int(gi["it0"]), int(gi["it1"])))
if rank == 0:
    print(f"world={world} cells/rank={mesh.n_cells} nr={nr} nl={nl} levels={g.amg_levels()} layout={g.amg_layout()} iters={its} res={gi['r1']:.2e} {time.time() - t0:.1f}s", flush=True)
if world > 1:
    dist.destroy_process_group()
