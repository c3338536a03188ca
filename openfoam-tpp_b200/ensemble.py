"""Parameter sweeps sharded over GPUs: one independent case per GPU, no data-path collective.

The reference builds sweeps from MATLAB-style ranges (main.py:118-142), zips the lists when they
all have the same length and takes their Cartesian product otherwise (main.py:504-534), then runs
the cases one after another locally (main.py:599-608) or one Slurm job each.  Here the cases of a
sweep are dealt round-robin to the ranks of a `torch.distributed` job (one process per GPU);
ranks exchange nothing but the final timing (max over ranks).
"""
from __future__ import annotations

import itertools


_SLACK = 1e-9  # inclusive-end tolerance of the reference's range syntax


def parse_range(text):
    """Sweep values from the reference's MATLAB-style syntax (main.py:118-142): `a:b` (step 1),
    `a:step:b`, or a comma list.  Same values as the reference produces: the k-th value is `a`
    plus k sequential additions of `step` (not a + k*step), the end is inclusive with a 1e-9
    slack, and range values are rounded to 6 decimals; list entries are taken as written."""
    import numpy as np

    fields = [f.strip() for f in text.strip().split(":")]
    if len(fields) == 1:
        return [float(tok) for tok in fields[0].split(",")]
    if len(fields) > 3:
        raise ValueError(f"Invalid range format: {text.strip()}")
    lo, hi = float(fields[0]), float(fields[-1])
    inc = float(fields[1]) if len(fields) == 3 else 1.0
    if not inc > 0:
        raise ValueError(f"Invalid range format: {text.strip()} (the step must be positive)")
    if lo > hi + _SLACK:
        return []
    room = int((hi - lo) / inc) + 3  # an upper bound on the count; the slack test below trims it
    running = np.cumsum(np.concatenate(([lo], np.full(room, inc))))  # sequential float64 additions
    keep = running <= hi + _SLACK
    return [round(float(v), 6) for v in running[: int(np.argmin(keep)) if not keep.all() else running.size]]


_NAME_FIELDS = (("H", "H"), ("D", "D"), ("", "geo"), ("R", "R"), ("f", "freq"), ("d", "duration"), ("m", "mesh"))


def case_name(params):
    """Directory name of a sweep member, the reference's naming contract (main.py:163-165):
    case_H<H>_D<D>_<geo>_R<R>_f<freq>_d<duration>_m<mesh>, values printed with str()."""
    return "_".join(["case"] + [tag + str(params[key]) for tag, key in _NAME_FIELDS])


def build_param_sets(base, sweeps):
    """Zip when every sweep list has the same length, Cartesian product otherwise
    (main.py:504-534, without the interactive confirmation)."""
    if not sweeps:
        return [dict(base)]
    keys = list(sweeps)
    lengths = {len(v) for v in sweeps.values()}
    combos = zip(*[sweeps[k] for k in keys]) if len(lengths) == 1 else itertools.product(*[sweeps[k] for k in keys])
    out = []
    for combo in combos:
        p = dict(base)
        p.update(dict(zip(keys, combo)))
        out.append(p)
    return out


def shard(items, world, rank):
    """Round-robin deal: rank r runs items r, r+world, ...  Every item lands on exactly one rank."""
    return list(items[rank::world])


def max_over_ranks(values, group=None):
    """Element-wise max of a list of floats over the ranks (the job's time is its slowest rank)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(values)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.tolist()


def sum_over_ranks(values, group=None):
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(values)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.tolist()


def run_sweep(root_dir, base, sweeps, lc_to_mesh=None, max_steps=None, lib_path=None, log=None, cases_per_gpu=1, write=True):
    """The sweep loop of main.py:599-608 on the GPUs of one box: every rank of the
    torch.distributed job (one process per GPU; a plain process is rank 0 of 1) sets up and runs
    its share of the cases, each through the ordinary case-directory path
    (case.setup_case -> foamrun.run_case).  Nothing is exchanged between ranks.

    cases_per_gpu > 1: that many cases of the rank's share advance at the same time, one host thread
    and one solver handle (CUDA stream) each ("one case per stream, 8 per GPU").
    The reference's cases are small (8 k - 42 k cells: a step is a chain of short kernels and a few
    host round trips), so one case alone leaves most of a B200 idle; the streams fill it.

    base / sweeps as in build_param_sets (keys H, D, geo, R, freq, duration, mesh = gmsh lc);
    lc_to_mesh(p) -> (n_rings, n_layers) replaces gmsh where it is absent (default: cells of
    about lc).  Returns [(case name, summary dict)] of this rank, in sweep order."""
    import os

    from . import case as cs
    from . import foamrun

    try:
        import torch.distributed as dist

        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    except ImportError:
        rank, world = 0, 1
    device = int(os.environ.get("LOCAL_RANK", "0"))
    lc_to_mesh = lc_to_mesh or (lambda p: (max(3, round(p["D"] / 2 / p["mesh"])), max(3, round(p["H"] / p["mesh"]))))
    mine = shard(build_param_sets(base, sweeps), world, rank)

    def one(p):
        name = case_name(p)
        d = os.path.join(root_dir, name)
        if not os.path.isdir(os.path.join(d, "constant", "polyMesh")):
            nr, nl = lc_to_mesh(p)
            cs.setup_case(d, H=p["H"], D=p["D"], geo=p["geo"], R=p["R"], freq=p["freq"], duration=p["duration"], n_rings=nr, n_layers=nl)
        out = foamrun.run_case(d, device=device, lib_path=lib_path, max_steps=max_steps, log=log, write=write)  # resumes from the latest time
        return name, out

    if cases_per_gpu <= 1 or len(mine) <= 1:
        return [one(p) for p in mine]
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=cases_per_gpu) as pool:
        return list(pool.map(one, mine))  # results in sweep order; an exception in any case propagates
