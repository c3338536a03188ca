"""Wall time per step and PCG iteration counts of the bench case for several TPP_* knob settings,
one mesh build for all of them (GPU):  python tools/knob_sweep.py "TPP_A=1 TPP_B=2" "TPP_A=2" ..."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from openfoam_tpp_b200 import solver as sv  # noqa: E402

cells = float(os.environ.get("SWEEP_CELLS", "6.2e6"))
mesh, nr, nl = bench.mesh_for(cells)
cfg = bench.make_config(mesh)
a0 = bench.initial_alpha(mesh)
for spec in sys.argv[1:] or [""]:
    kv = dict(x.split("=") for x in spec.split())
    for k, v in kv.items():
        os.environ[k] = v
    g = sv.Solver(mesh, cfg, device=0)
    g.set("alpha", a0)
    g.init_fields()
    g.step(4)
    t0 = time.perf_counter()
    its = []
    for _ in range(5):
        g.step(1)
        i = g.info()
        its.append((int(i["it0"]), int(i["it1"])))
    dt = (time.perf_counter() - t0) / 5
    print(f"{spec!r:60s} {dt * 1e3:7.2f} ms/step  {mesh.n_cells / dt / 1e6:6.1f} Mcell-steps/s  iters {its}", flush=True)
    g.close()
    for k in kv:
        os.environ.pop(k, None)
