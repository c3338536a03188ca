#!/usr/bin/env python
"""bench.py — whole-step throughput of the incompressibleVoF PIMPLE time step.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): the
D = 0.2 m, H = 0.208 m flat-bottom tank, orbital shaking R = 4 mm at 1.88 Hz
(case_H0.208_D0.2_flat_R0.004_f1.88_d20.0_m0.009), tet mesh synthetically refined to
--cells per GPU (default 6.2 M: the ~50 M-cell case of BASELINE.json over 8 GPUs).
A "step" is one full time step: Courant -> deltaT -> mesh motion -> 3 MULES sub-cycles ->
momentum assembly -> 2 pressure correctors (GAMG-PCG solves).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (see the task contract): value = cell-steps/s with the state resident
in HBM; e2e = the same metric through the C-ABI with pinned host buffers copied in and out
every step; roofline = the dominant kernel's algorithmic bytes / CUDA-event time against the
measured HBM peak; cpu_baseline = the CPU oracle (a restatement, not OpenFOAM) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cell_steps_per_s"
UNIT = "Mcell-steps/s"
CASE = dict(H=0.208, D=0.2, R=0.004, freq=1.88, duration=20.0, dt=0.001, ramp=2.0)


def mesh_for(cells_target):
    """n_rings / n_layers giving ~cells_target tets with near-isotropic cells."""
    from openfoam_tpp_b200 import meshgen

    # cells = 18 nr^2 nl ; isotropy: (D/2)/nr ~ H/nl  ->  nl = nr * H/(D/2)
    ratio = CASE["H"] / (CASE["D"] / 2)
    nr = max(4, int(round((cells_target / (18.0 * ratio)) ** (1.0 / 3.0))))
    nl = max(4, int(round(nr * ratio)))
    return meshgen.cylinder_mesh(CASE["H"], CASE["D"], nr, nl, "flat", "tet"), nr, nl


def make_config(mesh, freq=None):
    """The reference's numerics (case template) on the synthetic mesh, without touching disk."""
    import tempfile

    from openfoam_tpp_b200 import case as cs
    from openfoam_tpp_b200 import foamfile as ff
    from openfoam_tpp_b200 import motion

    with tempfile.TemporaryDirectory() as tmp:
        cs.write_template(tmp, end_time=CASE["duration"], fill_z=CASE["H"] / 2)
        rows = motion.orbital_table(CASE["R"], CASE["freq"] if freq is None else freq, 3.0, CASE["dt"], CASE["ramp"])
        motion.write_table(os.path.join(tmp, "constant", "6DoF.dat"), rows)
        cfg = cs.read_config(tmp, None)
        fields = {n: ff.read_field(os.path.join(tmp, "0", n)) for n in ("U", "alpha.water", "p_rgh")}
        cs._bc_tables(cfg, mesh, fields, "0")
    cfg.start_time = 0.0
    return cfg


def initial_alpha(mesh):
    from openfoam_tpp_b200 import meshgen

    C, _ = meshgen.cell_geometry(mesh)
    return (C[:, 2] <= CASE["H"] / 2).astype(np.float64)


VB = 4  # bytes per value inside the multigrid V-cycle (FP32 preconditioner; TPP_FP32=0 -> 8)
ALG_BYTES = {
    # algorithmic bytes per launch (SURVEY.md §8d convention: each distinct array once; FP64 8 B,
    # label 4 B; C cells, F internal faces).  V-cycle kernels (v_*) move VB-byte values.
    "v_jacobi": lambda C, F: 4 * VB * C + (VB + 8) * F,      # x, b, diag in; x out; upper + addressing
    "v_residual": lambda C, F: 4 * VB * C + (VB + 8) * F,
    "v_jacobi_first": lambda C, F: 3 * VB * C + (VB + 8) * F,        # b, diag in (own row and, re-read, the neighbours'); x out
    "v_jacobi_corr": lambda C, F: (4 * VB + 4) * C + (VB + 8) * F,   # x, b, diag, aggregate map in (+ the coarse vector, L2-resident); x out
    "v_spmv_dot2": lambda C, F: 4 * VB * C + (VB + 8) * F,   # c, r, diag in; A c out
    "spmv_dot": lambda C, F: 24 * C + 16 * F,
    "update_xr": lambda C, F: 48 * C,                       # x, r in/out; pA, wA in
    "update_p": lambda C, F: 24 * C,
    "reduce": lambda C, F: 16 * C,
    "init_residual": lambda C, F: 40 * C + 16 * F,
    "HbyA": lambda C, F: 88 * C + 16 * F,
    "Uf": lambda C, F: 24 * C + 88 * F,
    "flux": lambda C, F: 8 * C + 56 * F,
    "p_face": lambda C, F: 24 * C + 72 * F,
    "mom_cell": lambda C, F: 104 * C + 40 * F,
    "mules_phipsi": lambda C, F: 32 * F,
    "alphaphi_acc": lambda C, F: 24 * F,
    "grad_scalar": lambda C, F: 32 * C + 40 * F,
    "alpha_flux": lambda C, F: 56 * C + 56 * F,
    "mules_setup": lambda C, F: 40 * C + 32 * F,
    "mules_cell": lambda C, F: 16 * C + 32 * F,
    "mules_face": lambda C, F: 16 * C + 24 * F,
    "mules_face_final": lambda C, F: 16 * C + 56 * F,               # the face pass + phiBD in, alphaPhiUn out, alphaPhi in/out
    "mules_update": lambda C, F: 24 * C + 16 * F,
    "grad_U": lambda C, F: 96 * C + 40 * F,
    "mom_face": lambda C, F: 120 * C + 100 * F,
    "phiHbyA": lambda C, F: 100 * C + 120 * F,
    "p_cell": lambda C, F: 16 * C + 32 * F,
    "U_recon": lambda C, F: 56 * C + 48 * F,
}


class stdout_to_stderr:
    """fd-level redirect: NCCL prints its version banner on stdout at communicator creation, and
    the bench contract is ONE JSON line on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def _cpu_worker(q, barrier, cells_target, steps):
    """One CPU-oracle instance (its own process): build, 3 warm-up steps, then `steps` timed steps."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    mesh, nr, nl = mesh_for(cells_target)
    cfg = make_config(mesh)
    o = oracle.Oracle(mesh, cfg)
    o.set("alpha", initial_alpha(mesh))
    o.stage("alphaBCs")
    o.stage("mixture")
    o.step(3)
    barrier.wait()
    t0 = time.time()
    o.step(steps)
    t1 = time.time()
    q.put((mesh.n_cells, t0, t1))


def cpu_sample(cells_target, steps, procs=None):
    """The CPU oracle (serial FP64 restatement of the OpenFOAM algorithm; NOT OpenFOAM) on a
    reduced-resolution mesh of the same case.  The oracle is single-threaded, so the host's cores
    are used the only way it can use them: one independent instance per core, all stepping the same
    sample at the same time (the throughput an ideally scaling MPI run of that many cores would
    reach - memory-bandwidth contention between the instances included).  Returns Mcell-steps/s
    over the common window, cells of the sample mesh, seconds per step of one instance, processes."""
    import multiprocessing as mp

    procs = procs or max(1, len(os.sched_getaffinity(0)))
    ctx = mp.get_context("spawn")
    q, barrier = ctx.Queue(), ctx.Barrier(procs)
    ps = [ctx.Process(target=_cpu_worker, args=(q, barrier, cells_target, steps)) for _ in range(procs)]
    for p in ps:
        p.start()
    res = [q.get(timeout=600) for _ in ps]
    for p in ps:
        p.join()
    ncell = res[0][0]
    window = max(r[2] for r in res) - min(r[1] for r in res)
    return ncell * steps * procs / window / 1e6, ncell, sum(r[2] - r[1] for r in res) / procs / steps, procs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    val, ncell, sps, procs = cpu_sample(args.cpu_cells, max(1, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg4 D=0.2 H=0.208 flat tank, orbital 4 mm @ 1.88 Hz, tets; CPU sample mesh {ncell} cells"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"oracle (CPU restatement, not OpenFOAM: OpenFOAM 13 is not installed), {procs} independent single-threaded instances (one per host core) each stepping a {ncell}-cell mesh of the same case, {args.steps} steps after 3 warm-up, {sps * 1e3:.0f} ms/step per instance"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


SMALL = {
    # BASELINE.json configs 1-3 as concrete synthetic cases; meshes from the repo's
    # generators, sized like the reference's own (cfg1: 7 766 gmsh tets -> 7 776 here)
    "cfg1": dict(name="case_H0.004_D0.0221_flat_R0.005_f2.0", H=0.004, D=0.0221, geo="flat", R=0.005, freq=2.0, n_rings=12, n_layers=3),
    "cfg3": dict(name="case_H0.004_D0.0221_cap_R0.005_f2.0", H=0.004, D=0.0221, geo="cap", R=0.005, freq=2.0, n_rings=12, n_layers=3),
}


def small_config(which):
    """(mesh, config, initial alpha, workload text) of cfg1 / cfg2 / cfg3, through the ordinary case-directory
    path (template + motion table + mesh + setFields, then the case reader)."""
    import tempfile

    from openfoam_tpp_b200 import case as cs

    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, "case")
        if which == "cfg2":
            cs.setup_tutorial_case(d, nx=20, ny=40, nz=30)
            text = "cfg2 sloshingTank3D6DoF: closed tutorial tank, 20 x 40 x 30 hexes, tabulated 6-DoF motion (gen6DoF restated), p reference cell"
        else:
            c = SMALL[which]
            cs.setup_case(d, H=c["H"], D=c["D"], geo=c["geo"], R=c["R"], freq=c["freq"], duration=10.0, n_rings=c["n_rings"], n_layers=c["n_layers"])
            text = f"{which} {c['name']}: tets, n_rings {c['n_rings']}, n_layers {c['n_layers']}, orbital shaking {c['freq']} Hz (ramp 1 s)"
        case = cs.Case(d)
        a0 = case.fields["alpha.water"].internal_array(case.mesh.n_cells).copy()
    return case.mesh, case.cfg, a0, text + f", {case.mesh.n_cells} cells"


def run_sweep_config(args):
    """cfg5: the 64-case sweep of BASELINE.json (H x R x f grid in main.py's range syntax) sharded over the
    ranks, --cases-per-gpu cases at a time on every GPU (one host thread + one CUDA stream each).  value =
    all cell-steps of the sweep / the slowest rank's wall time; every case runs --steps steps."""
    import tempfile
    import time as _t

    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # more hardware queues than the default 8: one per concurrent case and copy stream
    import torch

    from openfoam_tpp_b200 import ensemble

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    base = {"H": 0.004, "D": 0.0221, "geo": "flat", "R": 0.005, "freq": 2.0, "duration": 10.0, "mesh": 0.0009}
    sweeps = {"H": ensemble.parse_range("0.004,0.008"), "R": ensemble.parse_range("0.002:0.001:0.005"), "freq": ensemble.parse_range("1.6:0.2:3.0")}
    lc_to_mesh = lambda p: (12, 3 if p["H"] < 0.006 else 6)  # 7 776 / 15 552 tets: the reference's small-case sizes
    sampler = ClockSampler(local)
    out = {}
    with tempfile.TemporaryDirectory() as tmp, stdout_to_stderr():
        # set-up (meshes, dictionaries, setFields) and a warm-up of --warmup steps are not timed
        ensemble.run_sweep(tmp, base, sweeps, lc_to_mesh=lc_to_mesh, max_steps=args.warmup, log=None, cases_per_gpu=args.cases_per_gpu, write=False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # (the one-case-at-a-time leg is a sample: the first 8 cases of the sweep, one value of R)
        seq_sweeps = dict(sweeps, R=sweeps["R"][:1])
        for mode, per in (("concurrent", args.cases_per_gpu), ("sequential", 1)):
            if mode == "concurrent":
                sampler.start()
            t0 = _t.perf_counter()
            done = ensemble.run_sweep(tmp, base, sweeps if mode == "concurrent" else seq_sweeps, lc_to_mesh=lc_to_mesh, max_steps=args.steps, log=None, cases_per_gpu=per, write=False)
            torch.cuda.synchronize()
            sec = _t.perf_counter() - t0
            if mode == "concurrent":
                sampler.stop_flag = True
            sec = ensemble.max_over_ranks([sec])[0]
            cs_ = ensemble.sum_over_ranks([float(sum(o["cells"] * o["steps"] for _, o in done)), float(sum(o["steps"] for _, o in done)), float(len(done))])
            out[mode] = {"seconds": sec, "cell_steps": cs_[0], "steps": cs_[1], "cases": int(cs_[2])}
    if rank == 0:
        c = out["concurrent"]
        line = {"metric": METRIC, "value": c["cell_steps"] / c["seconds"] / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": c["seconds"] / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"cfg5: {c['cases']} sweep cases (H 0.004,0.008 x R 0.002:0.001:0.005 x f 1.6:0.2:3.0, D 0.0221 flat, 7 776 / 15 552 tets), {args.steps} steps each after {args.warmup} warm-up steps, through foamrun.run_case (case directories, no field writes)",
                           "parallelism": f"ensemble: cases dealt round-robin to {world} GPU(s), {args.cases_per_gpu} concurrent cases (host threads / CUDA streams) per GPU, no collective",
                           "vof_steps_per_s": c["steps"] / c["seconds"], "vof_steps_per_s_one_case_at_a_time": out["sequential"]["steps"] / out["sequential"]["seconds"],
                           "concurrency_gain": (c["steps"] / c["seconds"]) / (out["sequential"]["steps"] / out["sequential"]["seconds"]), "sequential_sample_cases": out["sequential"]["cases"], "l2": "every case lives in L2 (a few MB): latency-bound, no HBM roofline (BASELINE.md section 4)"},
                "clocks": sampler.summary(), "gpu_launches": None, "e2e": {"value": c["cell_steps"] / c["seconds"] / 1e6, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "the sweep is timed end to end on the host clock (case directory in, solver stepping through the C-ABI)"},
                "roofline": None, "cpu_baseline": None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def parity_nranks(rank, world, local, dist, steps=3):
    """N > 1, before the timed loop (what `mpirun -np N foamRun -parallel` must guarantee,
    circularSloshingTank/Makefile:75-82): a ~100 k-cell tank decomposed into `world` z-slabs steps
    next to the same tank on one GPU (every rank runs its own copy), both solving p_rgh tightly.
    The time-step sequence must be identical and the fields agree (alpha 1e-9, U / p_rgh 1e-7 of
    scale).  Exercises the peer-memory halo / all-reduce kernels the timed run uses."""
    from openfoam_tpp_b200 import ensemble, meshgen
    from openfoam_tpp_b200 import solver as sv

    nr, nl = 14, max(world, 32 // world * world)
    whole = meshgen.cylinder_mesh(CASE["H"], CASE["D"], nr, nl, "flat", "tet")
    k0, k1 = rank * nl // world, (rank + 1) * nl // world
    part = meshgen.cylinder_mesh(CASE["H"], CASE["D"], nr, nl, "flat", "tet", k0=k0, k1=k1, proc=(rank, rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None))

    def tight(cfg):
        for sc in (cfg.p_rgh, cfg.p_rgh_final):
            sc.tolerance, sc.rel_tol, sc.max_iter = 1e-13, 0.0, 800
        return cfg

    g = sv.Solver(part, tight(make_config(part)), device=local)
    with stdout_to_stderr():
        g.comm_init_nccl()
    gw = sv.Solver(whole, tight(make_config(whole)), device=local)
    for s_, m_ in ((g, part), (gw, whole)):
        s_.set("alpha", initial_alpha(m_))
        s_.init_fields()
    cpl = whole.n_cells // nl
    sl = slice(k0 * cpl, k1 * cpl)
    errs = {"alpha": 0.0, "U": 0.0, "p_rgh": 0.0}
    dt_equal = True
    for _ in range(steps):
        g.step(1)
        gw.step(1)
        gi, wi = g.info(), gw.info()
        dt_equal = dt_equal and abs(gi["t"] - wi["t"]) <= 1e-12 * wi["t"]
        for nm, nc in (("alpha", 1), ("U", 3), ("p_rgh", 1)):
            a, b = g.get(nm), gw.get(nm).reshape(-1, nc)
            errs[nm] = max(errs[nm], float(np.abs(a - b[sl].reshape(-1)).max() / max(np.abs(b).max(), 1e-300)))
    worst = ensemble.max_over_ranks([errs["alpha"], errs["U"], errs["p_rgh"], 0.0 if dt_equal else 1.0])
    g.close()
    gw.close()
    ok = worst[0] <= 1e-9 and worst[1] <= 1e-7 and worst[2] <= 1e-7 and worst[3] == 0.0
    return {"ok": bool(ok), "cells": whole.n_cells, "ranks": world, "steps": steps, "time_step_sequence_equal": worst[3] == 0.0,
            "max_rel_err": {"alpha": worst[0], "U": worst[1], "p_rgh": worst[2]}, "tolerance": {"alpha": 1e-9, "U": 1e-7, "p_rgh": 1e-7}}


def main():
    global VB
    VB = 8 if os.environ.get("TPP_FP32", "1") == "0" else 4
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cells", type=float, default=6.2e6, help="target cells per GPU (6.2 M = the ~50 M-cell tank of BASELINE config 4 over 8 GPUs)")
    ap.add_argument("--cpu-cells", type=float, default=1.2e5, help="cells of the CPU-baseline sample mesh")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kernel-table", default=None, help="write the full per-kernel table (launches, ms, algorithmic GB/s) of the profiled steps to this JSON file")
    ap.add_argument("--mode", default="decomposed", choices=["decomposed", "ensemble"], help="N > 1: one tank over N GPUs (halo exchange) or N independent sweep cases")
    ap.add_argument("--config", default="cfg4", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs: cfg4 (default) = the refined D 0.2 m tank, the configuration the metric is quoted on; cfg1 / cfg3 = the small flat / cap cylinders, cfg2 = the sloshingTank3D6DoF tutorial tank, cfg5 = the 64-case sweep (cases sharded over the GPUs, --cases-per-gpu at a time on each)")
    ap.add_argument("--cases-per-gpu", type=int, default=8)
    ap.add_argument("--t0", type=float, default=None, help="start time of the run: 2.0 = the end of the shaker's ramp (full orbit amplitude, SURVEY.md 8d); 0 = from rest")
    ap.add_argument("--spinup", type=int, default=45, help="untimed steps before the --warmup steps (start-up transient of the impulsively started tank)")
    ap.add_argument("--order", default="structured", choices=["structured", "gmsh"], help="cell order of the mesh file: the generator's (layer by layer) or a random one, as a gmsh mesh has")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the decomposed-vs-whole check before the timed loop")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    if args.t0 is None:
        args.t0 = 2.0 if args.config == "cfg4" else 0.0
    if args.config == "cfg5":
        return run_sweep_config(args)

    import torch

    from openfoam_tpp_b200 import solver as sv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()

    from openfoam_tpp_b200 import ensemble, meshgen

    freqs = [CASE["freq"]]
    if world > 1 and args.mode == "ensemble":
        # the sweep of BASELINE.json config 5 sharded one independent case per GPU (the same tank
        # at a different shaking frequency on every rank); no data-path collective
        freqs = ensemble.parse_range(f"{CASE['freq']}:0.04:{CASE['freq'] + 0.04 * (world - 1) + 1e-6}")
        mesh, nr, nl = mesh_for(args.cells)
        cfg = make_config(mesh, ensemble.shard(freqs, world, rank)[0])
    elif world > 1:
        # ONE tank decomposed over the GPUs (BASELINE.json config 4): the mesh is refined so that
        # every GPU keeps --cells cells (weak scaling), cut into z-slabs (decomposePar `simple`
        # n = (1 1 N)); each rank builds only its slab, the cuts are `processor` patches
        ratio = CASE["H"] / (CASE["D"] / 2)
        nr = max(4, int(round((args.cells * world / (18.0 * ratio)) ** (1.0 / 3.0))))
        nl = max(world, int(round(nr * ratio / world)) * world)
        k0, k1 = rank * nl // world, (rank + 1) * nl // world
        mesh = meshgen.cylinder_mesh(CASE["H"], CASE["D"], nr, nl, "flat", "tet", k0=k0, k1=k1,
                                     proc=(rank, rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None))
        cfg = make_config(mesh)
    elif args.config != "cfg4":
        mesh, cfg, a0, workload = small_config(args.config)
        nr = nl = 0
    else:
        mesh, nr, nl = mesh_for(args.cells)
        cfg = make_config(mesh)
    if args.order == "gmsh":
        mesh = meshgen.shuffled(mesh, seed=1234 + rank)
    nC, nI, nF = mesh.n_cells, mesh.n_internal, mesh.n_faces
    decomposed = world > 1 and args.mode != "ensemble"
    parity = None
    if decomposed and not args.no_parity:
        parity = parity_nranks(rank, world, local, dist)
    t_create = time.perf_counter()
    g = sv.Solver(mesh, cfg, device=local)
    t_create = time.perf_counter() - t_create  # tpp_create: addressing, geometry, renumbering decision, uploads
    # a side stream shared with the solver: CUDA events recorded here bracket its kernels, and
    # (unlike the legacy default stream) it can be graph-captured
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g.use_stream(stream.cuda_stream)
    if decomposed:
        with stdout_to_stderr():
            g.comm_init_nccl()
            torch.cuda.synchronize()
    if args.config == "cfg4":
        a0 = initial_alpha(mesh)
        workload = f"cfg4 case_H0.208_D0.2_flat_R0.004_f1.88 tet mesh refined to {nC} cells per GPU (n_rings {nr}, n_layers {nl})"
    g.set("alpha", a0)
    g.init_fields()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-state throughput ---------------------------------------------------------
    if args.t0 > 0:
        g.set_time(args.t0, cfg.delta_t)  # the shaker at full orbit amplitude (end of its ramp), fluid still at rest
    t_first = time.perf_counter()
    g.step(1)  # the first step builds the multigrid hierarchy (device matching, host coarse graphs; cached afterwards)
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t_first
    g.step(max(args.spinup - 1, 0))
    g.step(args.warmup)
    g.stats(1)
    l0 = g.info()["launches"]
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    g.step(args.steps)
    e1.record(stream)
    barrier()
    sec = e0.elapsed_time(e1) / 1e3
    sampler.stop_flag = True
    info = g.info()
    stats = g.stats(0)
    launches = int(info["launches"] - l0) - 4 * args.steps  # (the statistics' own two small reductions per step are not counted)

    # ---- per-kernel CUDA-event timing (separate pass, same state) ---------------------------
    g.profile(True)
    g.step(2)
    prof = g.profile_report()
    g.profile(False)
    tot_ms = sum(v[1] for v in prof.values())
    top = sorted(prof.items(), key=lambda kv: -kv[1][1])
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    levels = g.amg_levels()
    layout = g.amg_layout()
    coarse = levels[1:layout["kernel_levels"]]  # CSR levels smoothed kernel by kernel (the rest live in the tail kernel)

    klev = levels[: max(layout["kernel_levels"], 1)]  # mesh level + the coarse levels smoothed kernel by kernel

    def alg_bytes(name, launches):
        """algorithmic bytes of ALL launches of a kernel in the profiled window"""
        if name in ALG_BYTES:
            return ALG_BYTES[name](nC, mesh.n_internal) * launches
        per_row = {"v_jacobi_csr": 4 * VB, "v_residual_csr": 4 * VB, "v_spmv_dot2_csr": 4 * VB, "v_jacobi_first_csr": 3 * VB, "v_jacobi_corr_csr": 4 * VB + 4}.get(name)
        if per_row and coarse:  # one launch per coarse level and sweep: bytes summed over the levels
            sweep = sum(per_row * n + (VB + 8) * f for n, f in coarse)
            return sweep * launches / len(coarse)
        # streaming V-cycle kernels that run once per kernel level: bytes per row, summed over those levels
        per_row = {"v_scale_apply": 6 * VB, "v_jacobi0": 3 * VB, "v_prolong": 4 + 2 * VB, "v_restrict": 4 + 2 * VB}.get(name)
        if per_row:
            return sum(per_row * n for n, _ in klev) * launches / len(klev)
        return None

    rated = [(k, v) for k, v in top if alg_bytes(k, v[0])]
    dom, (dom_n, dom_ms) = rated[0] if rated else top[0]
    ab = alg_bytes(dom, dom_n)
    achieved = ab / (dom_ms * 1e-3) / 1e9 if ab else None
    table = []
    for k, v in top[:10]:
        b = alg_bytes(k, v[0])
        table.append([k, v[0], round(v[1], 3), round(b / (v[1] * 1e-3) / 1e9, 1) if b else None])
    if args.kernel_table and rank == 0:
        full = [[k, v[0], round(v[1], 4), (round(alg_bytes(k, v[0]) / (v[1] * 1e-3) / 1e9, 1) if alg_bytes(k, v[0]) else None)] for k, v in top]
        json.dump({"profiled_steps": 2, "cells": nC, "internal_faces": nI, "total_ms": tot_ms, "kernels_launches_ms_GBps": full}, open(args.kernel_table, "w"), indent=1)
    # measured DRAM traffic of the dominant kernel (committed ncu --set full capture), scaled per cell
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")
    if os.path.exists(tp):
        tr = json.load(open(tp))
        if dom in tr["kernels"] and "dram_bytes" in tr["kernels"][dom]:
            traffic = tr["kernels"][dom]["dram_bytes"] * nC / tr["cells"]
    # the whole step against the roofline: algorithmic bytes of every kernel with a byte model
    # (measured launch counts, i.e. measured iteration counts) over the graph-mode step time
    step_bytes = sum(alg_bytes(k, v[0]) or 0 for k, v in top) / 2
    modelled_ms = sum(v[1] for k, v in top if alg_bytes(k, v[0]))
    step_roof = {"algorithmic_bytes_per_step": step_bytes, "achieved": step_bytes / (sec / args.steps) / 1e9, "frac": step_bytes / (sec / args.steps) / 1e9 / peak,
                 "modelled_share_of_kernel_time": modelled_ms / tot_ms, "bytes_per_cell_step": step_bytes / nC}
    roofline = {"bound": "hbm", "kernel": f"k_{dom}", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
                "traffic": traffic, "algorithmic_bytes": ab / dom_n if ab else None, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "share_of_step": dom_ms / tot_ms, "step": step_roof, "amg_levels_rows_faces": levels, "amg_layout": layout,
                "top_kernels_launches_ms_GBps": table}

    # ---- end to end: pinned host state in, one step, host state out, every step ---------------
    # The state a restart needs goes in (alpha, U, p_rgh, phi, Uf) and the written fields plus that
    # state come out (alpha, U, p_rgh, p, phi, Uf), all through tpp_set / tpp_get on pinned host
    # buffers; the next step's input is this step's output (same buffers, no host-side copy).
    names_in = ["alpha", "U", "p_rgh", "phi", "Uf"]
    names_out = ["alpha", "U", "p_rgh", "p", "phi", "Uf"]
    host = {n: torch.from_numpy(g.get(n)).pin_memory() for n in names_out}
    import ctypes as C

    from openfoam_tpp_b200 import abi

    def dp(tn):
        return C.cast(tn.data_ptr(), abi.c_double_p)

    h2d_b = sum(host[n].numel() * 8 for n in names_in)
    d2h_b = sum(host[n].numel() * 8 for n in names_out)
    # The next step's input is this step's output (same pinned buffers, no host-side copy), so the
    # three phases of a step cannot overlap: input copy, step, output copy are timed back to back.
    # (tpp_get_async streams results out behind the next step - 55 -> 46 ms per step when the inputs do
    # not depend on them, tools/e2e_probe.py - but feeding a stale state back in is not a simulation.)
    e2e_steps = max(2, args.steps // 2)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for n in names_in:
            g.L.tpp_set(g.h, n.encode(), dp(host[n]), host[n].numel())
        g.step(1)
        for n in names_out:
            g.L.tpp_get(g.h, n.encode(), dp(host[n]), host[n].numel())
    barrier()
    sec_e2e = time.perf_counter() - t0

    sec, sec_e2e, it0_mean, it1_mean, it1_max, cap_hits, bal = ensemble.max_over_ranks([sec, sec_e2e, stats["it0_mean"], stats["it1_mean"], stats["it1_max"], stats["cap_hits"], 0.0])
    vols = ensemble.sum_over_ranks([stats["alpha_volume_start"], stats["alpha_volume"], stats["alpha_boundary_outflow"]])
    bal = abs(vols[1] - vols[0] + vols[2]) / max(abs(vols[0]), 1e-300)
    checks = {"alpha_volume_balance_rel": bal, "alpha_volume_balance_ok": bool(bal <= 1e-10), "p_rghFinal_max_iter": int(cfg.p_rgh_final.max_iter),
              "p_rghFinal_iters_max": int(it1_max), "p_rghFinal_below_max_iter": bool(it1_max < cfg.p_rgh_final.max_iter), "steps_capped": int(cap_hits)}
    total_cells = int(ensemble.sum_over_ranks([float(nC)])[0])
    launches = int(ensemble.sum_over_ranks([float(launches)])[0])
    value = total_cells * args.steps / sec / 1e6
    e2e = total_cells * e2e_steps / sec_e2e / 1e6

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, ncell, sps, procs = cpu_sample(args.cpu_cells, args.cpu_steps)
        cpu = {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
               "sample": f"CPU oracle (restatement, not OpenFOAM), {procs} independent single-threaded instances (one per host core) each stepping a {ncell}-cell mesh of the same case, {args.cpu_steps} steps after 3 warm-up, {sps * 1e3:.0f} ms/step per instance"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "launches_per_step": launches / max(args.steps, 1) / world,
                       "cells_per_gpu": nC, "internal_faces_per_gpu": nI, "parallelism": "single GPU" if world == 1 else (f"one tank decomposed into {world} z-slabs (simple (1 1 {world})), NCCL halo exchange + all-reduced Krylov dots, {total_cells} cells in total" if decomposed else f"ensemble: {world} independent sweep cases (f = {freqs[0]}..{freqs[-1]} Hz), one per GPU, no collective"),
                       "l2": "working set (>1 kB/cell) far exceeds the 126 MB L2; no flush needed",
                       "vof_steps_per_s": args.steps / sec, "solver_iters_last_step": [int(info["it0"]), int(info["it1"])], "amg_levels": int(info["levels"]),
                       "t_start": args.t0, "t_end": float(info["t"]), "spinup_steps": args.spinup, "iters_mean": [round(it0_mean, 2), round(it1_mean, 2)],
                       "cell_order": args.order, "setup_seconds": {"tpp_create": round(t_create, 2), "first_step_with_multigrid_build": round(t_first, 2)},
                       "precision": "FP64 fields, operators, Krylov iteration and residuals; multigrid preconditioner in " + ("FP64" if VB == 8 else "FP32")},
            "clocks": sampler.summary(), "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b, "steps": e2e_steps,
                    "how": "per step: tpp_set of the restart state from pinned host memory, tpp_step, tpp_get of the written fields + that state into the same pinned buffers (the next step's input is this step's output)"},
            "roofline": roofline, "cpu_baseline": cpu, "checks": checks,
        }
        if parity is not None:
            line["parity_nranks"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
