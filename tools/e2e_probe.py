"""Where the end-to-end step goes: H2D of the restart state, the step, blocking and streamed D2H
(tuning aid for bench.py's e2e loop).  gpurun -- python tools/e2e_probe.py"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = ["x"]
import torch  # noqa: E402

import bench  # noqa: E402
from openfoam_tpp_b200 import abi  # noqa: E402
from openfoam_tpp_b200 import solver as sv  # noqa: E402

mesh, nr, nl = bench.mesh_for(6.2e6)
g = sv.Solver(mesh, bench.make_config(mesh))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
g.use_stream(stream.cuda_stream)
g.set("alpha", bench.initial_alpha(mesh))
g.init_fields()
g.step(5)
names_in = ["alpha", "U", "p_rgh", "phi", "Uf"]
names_out = ["alpha", "U", "p_rgh", "p", "phi", "Uf"]
host = {n: torch.from_numpy(g.get(n)).pin_memory() for n in names_out}
out = {n: torch.empty_like(host[n]).pin_memory() for n in names_out}
print("pinned:", all(t.is_pinned() for t in host.values()), all(t.is_pinned() for t in out.values()))
dp = lambda t: C.cast(t.data_ptr(), abi.c_double_p)


def timed(label, fn, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    g.L.tpp_sync(g.h)
    torch.cuda.synchronize()
    print(f"{label:34s} {(time.perf_counter() - t0) / reps * 1e3:8.2f} ms")


def h2d():
    for n in names_in:
        g.L.tpp_set(g.h, n.encode(), dp(host[n]), host[n].numel())


def d2h_block():
    for n in names_out:
        g.L.tpp_get(g.h, n.encode(), dp(out[n]), out[n].numel())


def d2h_async():
    for n in names_out:
        g.L.tpp_get_async(g.h, n.encode(), dp(out[n]), out[n].numel())


d2h_async()
g.L.tpp_sync(g.h)
timed("tpp_set x5 (H2D 647 MB)", h2d)
timed("tpp_step", lambda: g.step(1))
timed("tpp_get x6 (D2H 697 MB, blocking)", d2h_block)
timed("tpp_get_async x6 + tpp_sync", lambda: (d2h_async(), g.L.tpp_sync(g.h)))
timed("set + step + get (blocking)", lambda: (h2d(), g.step(1), d2h_block()))
timed("set + step + get_async (streamed)", lambda: (h2d(), g.step(1), d2h_async()), reps=8)
