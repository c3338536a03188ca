"""`foamRun` for the incompressibleVoF module, on the GPU.

Drop-in for the two ways the reference launches its solver:
  * main.py:333-348  run_case_local(case_dir, n_cpus) -> make -C case run|resume
  * circularSloshingTank/Makefile:85,98  `foamRun` with cwd = case directory
    (`python -m openfoam_tpp_b200.foamrun [-case DIR] [-parallel]` accepts foamRun's argv)

Reads the case (constant/polyMesh, constant/*, system/*, latest time directory), steps it with
libtppvof.so through the C-ABI, and writes what OpenFOAM would: binary time directories named
by `timeFormat general; timePrecision 6` holding alpha.water, U, p_rgh, p, rho, phi, Uf,
the moved polyMesh/points and uniform/time, and postProcessing/probes/<start>/p in OpenFOAM's
text layout (reference artefact: case_H0.004_D0.0221_flat_R0.005_f2.0/postProcessing/probes/0/p).
"""
from __future__ import annotations

import os
import sys
import time as _time

import numpy as np

from . import foamfile as ff
from .case import Case
from .foamfile import FoamError
from .solver import Solver


def _g(v):
    return f"{v:.6g}"


def _boundary_dict(case, name, values_by_patch=None):
    """boundaryField entries: types carried over from the start field, fresh values."""
    out = {}
    src = case.fields[name].boundary if name in case.fields else {}
    for p in case.mesh.patches:
        e = {}
        s = src.get(p["name"], {"type": "calculated"})
        for k, v in s.items():
            if k != "value":
                e[k] = v
        if values_by_patch is not None and p["name"] in values_by_patch:
            e["value"] = values_by_patch[p["name"]]
        out[p["name"]] = e
    return out


def write_time(case, solver, name, binary=True, precision=6):
    """One OpenFOAM time directory from the device state."""
    mesh = case.mesh
    nI = mesh.n_internal
    tdir = os.path.join(case.dir, name)

    nBphys = sum(q["nFaces"] for q in mesh.patches if q["type"] != "processor")  # the solver's boundary arrays hold these

    def patch_slices(arr, ncomp):
        """per-patch views of an array over ALL boundary faces of the file (surface fields)"""
        a = arr.reshape(-1, ncomp) if ncomp > 1 else arr
        return {p["name"]: a[p["startFace"] - nI : p["startFace"] - nI + p["nFaces"]] for p in mesh.patches}

    def vol_patches(cells, bnd, ncomp):
        """vol field: physical patches from the solver's boundary array, processor patches (not
        needed by a restart or by reconstructPar) from the adjacent cells"""
        b = bnd.reshape(-1, ncomp) if ncomp > 1 else bnd
        out = {}
        for p in mesh.patches:
            sl = slice(p["startFace"], p["startFace"] + p["nFaces"])
            out[p["name"]] = cells[mesh.owner[sl]] if p["type"] == "processor" else b[p["startFace"] - nI : p["startFace"] - nI + p["nFaces"]]
        return out

    alpha, alpha_b = solver.get("alpha"), solver.get("alpha_b")
    U, U_b = solver.get("U").reshape(-1, 3), solver.get("U_b")
    p_rgh, p_rgh_b = solver.get("p_rgh"), solver.get("p_rgh_b")
    rho, rho_b = solver.get("rho"), solver.get("rho_b")
    p = solver.get("p")
    phi = solver.get("phi")
    Uf = solver.get("Uf").reshape(-1, 3)
    W = lambda fld: ff.write_field(os.path.join(tdir, fld.name), fld, binary=binary, precision=precision, location=name)
    W(ff.Field("volScalarField", "alpha.water", "[0 0 0 0 0 0 0]", alpha, _boundary_dict(case, "alpha.water", vol_patches(alpha, alpha_b, 1))))
    W(ff.Field("volVectorField", "U", "[0 1 -1 0 0 0 0]", U, _boundary_dict(case, "U", vol_patches(U, U_b, 3))))
    W(ff.Field("volScalarField", "p_rgh", "[1 -1 -2 0 0 0 0]", p_rgh, _boundary_dict(case, "p_rgh", vol_patches(p_rgh, p_rgh_b, 1))))
    calc = {q["name"]: {"type": "processor" if q["type"] == "processor" else "calculated"} for q in mesh.patches}
    pb = p_rgh_b + rho_b * solver.get("ghf")[nI : nI + nBphys]
    bd = {k: dict(v, value=vol_patches(p, pb, 1)[k]) for k, v in calc.items()}
    W(ff.Field("volScalarField", "p", "[1 -1 -2 0 0 0 0]", p, bd))
    bd = {k: dict(v, value=vol_patches(rho, rho_b, 1)[k]) for k, v in calc.items()}
    W(ff.Field("volScalarField", "rho", "[1 -3 0 0 0 0 0]", rho, bd))
    bd = {k: dict(v, value=patch_slices(phi[nI:], 1)[k]) for k, v in calc.items()}
    W(ff.Field("surfaceScalarField", "phi", "[0 3 -1 0 0 0 0]", phi[:nI], bd))
    bd = {k: dict(v, value=patch_slices(Uf[nI:].reshape(-1), 3)[k]) for k, v in calc.items()}
    W(ff.Field("surfaceVectorField", "Uf", "[0 1 -1 0 0 0 0]", Uf[:nI], bd))
    if case.cfg.motion is not None:
        os.makedirs(os.path.join(tdir, "polyMesh"), exist_ok=True)
        ff.write_points(os.path.join(tdir, "polyMesh", "points"), solver.get("points").reshape(-1, 3), binary, f"{name}/polyMesh")
    info = solver.info()
    os.makedirs(os.path.join(tdir, "uniform"), exist_ok=True)
    with open(os.path.join(tdir, "uniform", "time"), "w") as f:
        f.write(ff._hdr("dictionary", "time", f"{name}/uniform"))
        f.write(f"value           {float(info['t'])!r};\n\nname            \"{name}\";\n\nindex           {int(info['step'])};\n\ndeltaT          {float(info['dt'])!r};\n\ndeltaT0         {float(info['dt'])!r};\n")
        f.write(ff.END)


class ProbesWriter:
    """postProcessing/probes/<startTime>/<field> in OpenFOAM's layout."""

    def __init__(self, case, start_name, field="p"):
        d = os.path.join(case.dir, "postProcessing", "probes", start_name)
        os.makedirs(d, exist_ok=True)
        self.f = open(os.path.join(d, field), "w")
        locs = case.cfg.probes
        for i, x in enumerate(locs):
            self.f.write(f"# Probe {i} ({_g(x[0])} {_g(x[1])} {_g(x[2])})\n")
        self.f.write(f"{'# Time':<13} " + " ".join(f"{i:<13}" for i in range(len(locs))) + "\n")

    def rows(self, arr):
        for r in arr:
            self.f.write(f"{_g(r[0]):<13} " + " ".join(f"{_g(v):<13}" for v in r[1:]).rstrip(" ") + "\n")
        self.f.flush()

    def close(self):
        self.f.close()


def run_case(case_dir, device=0, lib_path=None, max_steps=None, log=sys.stdout, write=True, parallel=False, interface=None):
    """Advance a case from its latest time to endTime.  Returns a summary dict.

    parallel: `foamRun -parallel` - this process is one rank of a torch.distributed job
    (torchrun / mpirun-style RANK, WORLD_SIZE); it runs the `processor<rank>` share written by
    decomposePar (the reference's Makefile:77-78, or decompose.decompose_par here) and writes its
    time directories there, for reconstructPar."""
    rank, world = 0, 1
    if parallel:
        import torch.distributed as dist

        if not dist.is_initialized():
            import torch

            nccl = torch.cuda.is_available() and lib_path is None
            if nccl:
                device = int(os.environ.get("LOCAL_RANK", "0"))
                torch.cuda.set_device(device)
                dist.init_process_group("nccl", device_id=torch.device("cuda", device))
            else:
                dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        if rank != 0:
            log = None
    case = Case(case_dir, processor=rank if parallel else None)
    cfg = case.cfg
    if parallel and world > 1:
        # every rank resolved `latestTime` in its own processorN/: they must agree
        from . import ensemble

        hi = ensemble.max_over_ranks([case.start_value, -case.start_value])
        if hi[0] != case.start_value or -hi[1] != case.start_value:
            raise FoamError(f"{case.dir}: start time {case.start_name} differs between the processor directories (latest complete times {-hi[1]:g} .. {hi[0]:g}); remove the partial time directories")
    s = Solver(case.mesh, cfg, device=device, lib_path=lib_path)
    if parallel and world > 1:
        if lib_path is None:
            s.comm_init_nccl()
        else:
            s.comm_init_callbacks()
    s.load_case_fields(case)
    nF = case.mesh.n_faces
    if "phi" in case.fields and "Uf" in case.fields and case.start_value > 0:
        # restart: internal + boundary values of the face fields
        def full(fld, ncomp):
            nI = case.mesh.n_internal
            a = np.zeros((nF, ncomp)) if ncomp > 1 else np.zeros(nF)
            a[:nI] = fld.internal_array(nI)
            for p in case.mesh.patches:
                v = fld.boundary.get(p["name"], {}).get("value")
                if v is not None:
                    a[p["startFace"] : p["startFace"] + p["nFaces"]] = v
            return a

        s.set("phi", full(case.fields["phi"], 1))
        s.set("Uf", full(case.fields["Uf"], 3))
        s.set_time(case.start_value, case.restart_delta_t or cfg.delta_t)
    probes = None
    multi = parallel and world > 1
    vg = -1.79769e307

    def merge_rows(rows):
        """probe rows of a decomposed run: every probe lives on one rank (-1.79769e307 elsewhere,
        OpenFOAM's own 'not found' value), so the element-wise maximum over the ranks is the row"""
        a = np.ascontiguousarray(np.asarray(rows, dtype=np.float64).reshape(-1, 1 + len(cfg.probes)))
        if multi:
            import torch
            import torch.distributed as dist

            t = torch.from_numpy(a.copy())
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            a = t.cpu().numpy()
        return a

    if cfg.probes is not None and len(cfg.probes):
        cells = [s.find_cell(x) for x in cfg.probes]
        s.set_probes(cells)
        if rank == 0:  # the master writes postProcessing/ at the case root, as OpenFOAM does
            root_case = case if not multi else type("RootDir", (), {"dir": case.root, "cfg": cfg})()
            probes = ProbesWriter(root_case, case.start_name, cfg.probe_fields[0] if cfg.probe_fields else "p")
        pnow = s.get("p")
        row0 = merge_rows([[case.start_value] + [pnow[c] if c >= 0 else vg for c in cells]])
        if probes is not None:
            probes.rows(row0)
    # in-situ interface statistics (opt-in: `interface=True`, `foamRun -interface` or TPP_INTERFACE=1):
    # the rows the reference's extract_interface derives from the time directories afterwards
    # (main.py:751-780), computed on the device at every write time
    if interface is None:
        interface = os.environ.get("TPP_INTERFACE", "0") == "1"
    iface = None
    if interface and not multi:
        os.makedirs(os.path.join(case.dir, "postProcessing", "interface"), exist_ok=True)
        iface = open(os.path.join(case.dir, "postProcessing", "interface", "interface_summary.csv"), "a" if case.start_value > 0 else "w")
        if case.start_value == 0:
            iface.write("time,max_z,min_z,mean_z,num_points")
            iface.write("\n{},{},{},{},{}".format(*s.interface_summary()))
    t0 = _time.perf_counter()
    steps0 = s.info()["step"]
    n_writes = 0
    while True:
        budget = 10**9 if max_steps is None else max(0, max_steps - int(s.info()["step"] - steps0))
        if budget == 0:
            break
        rc = s.run_to_write(budget)
        info = s.info()
        if cfg.probes is not None and len(cfg.probes):
            rows = merge_rows(s.probe_log())
            if probes is not None:
                probes.rows(rows)
        if rc == 1:
            name = ff.time_name(info["t"], cfg.time_precision)
            if write:
                write_time(case, s, name, cfg.write_binary, cfg.write_precision)
            n_writes += 1
            if iface is not None:
                iface.write("\n{},{},{},{},{}".format(*s.interface_summary()))
                iface.flush()
            if log:
                el = _time.perf_counter() - t0
                print(f"Time = {name}  step {int(info['step'])}  deltaT = {info['dt']:.6g}  Co = {info['Co']:.3g}  p_rghFinal iters {int(info['it1'])} res {info['r1']:.2e}  ExecutionTime = {el:.2f} s", file=log, flush=True)
        else:
            break
    el = _time.perf_counter() - t0
    info = s.info()
    nsteps = int(info["step"] - steps0)
    if probes is not None:
        probes.close()
    if iface is not None:
        iface.close()
    ncells = case.mesh.n_cells
    if parallel and world > 1:
        from . import ensemble

        ncells = int(ensemble.sum_over_ranks([float(ncells)])[0])
    out = {"steps": nsteps, "seconds": el, "t": info["t"], "writes": n_writes, "cells": ncells, "mcell_steps_per_s": ncells * nsteps / max(el, 1e-30) / 1e6}
    s.close()
    return out


def run_case_local(case_dir, n_cpus=1, device=0):
    """Signature-compatible replacement of main.py:333-348 (n_cpus is accepted and ignored: one
    GPU runs the case; resume-vs-run is decided by the latest time directory, as `startFrom
    latestTime` does in the reference's controlDict:19)."""
    print(f"Running case {case_dir} on GPU {device}...")
    return run_case(case_dir, device=device)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    case_dir = os.getcwd()
    parallel = False
    while argv:
        a = argv.pop(0)
        if a == "-case":
            case_dir = argv.pop(0)
        elif a == "-parallel":
            # one rank of `torchrun --nproc-per-node N -m openfoam_tpp_b200.foamrun -parallel`; without
            # a launcher (WORLD_SIZE unset) the whole case runs on one GPU, as before
            parallel = int(os.environ.get("WORLD_SIZE", "1")) > 1
        elif a == "-noFunctionObjects":
            pass
        elif a == "-interface":
            os.environ["TPP_INTERFACE"] = "1"
        elif a == "-solver":
            if argv.pop(0) != "incompressibleVoF":
                raise SystemExit("only the incompressibleVoF solver module is provided")
        else:
            raise SystemExit(f"foamRun (tppvof): unknown option {a}")
    try:
        out = run_case(case_dir, parallel=parallel)
    except Exception as e:  # non-zero exit status, as `check=True` expects (main.py:345,348)
        print(f"--> FOAM FATAL ERROR: {e}", file=sys.stderr)
        return 1
    print(f"End  ({out['steps']} steps, {out['seconds']:.1f} s, {out['mcell_steps_per_s']:.2f} Mcell-steps/s)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
