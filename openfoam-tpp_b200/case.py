"""Case directories: the contract between the reference's orchestrator and the solver.

`read_case` turns what `main.py:setup_case` (main.py:266-331) + `gmshToFoam` + `setFields`
(circularSloshingTank/Makefile:73-74) leave on disk into a `CaseConfig`; unknown schemes,
solvers or boundary conditions are hard errors naming file and keyword (no
silent defaults).  `setup_case` is the stand-in for the parts of that pipeline that need
tools absent from this image (gmsh, gmshToFoam, setFields): it writes the same dictionaries
with the same values, generates constant/polyMesh with the repo's own mesher and applies
the box initialisation of system/setFieldsDict.
"""
from __future__ import annotations

import os
import shutil
from dataclasses import dataclass, field

import numpy as np

from . import foamfile as ff
from . import meshgen, motion
from .foamfile import FoamError, lookup, to_bool, to_float, to_vector

U_BC = {"movingWallVelocity": 0, "pressureInletOutletVelocity": 1}
A_BC = {"zeroGradient": 0, "inletOutlet": 1}
P_BC = {"fixedFluxPressure": 0, "totalPressure": 1}
SMOOTHERS = {"DIC": 0, "DICGaussSeidel": 1, "GaussSeidel": 2}


@dataclass
class SolverControl:
    type: int = 0  # 0 PCG, 1 GAMG
    precond: int = 0  # PCG: 0 DIC, 1 GAMG
    smoother: int = 0
    tolerance: float = 1e-6
    rel_tol: float = 0.0
    max_iter: int = 1000
    n_vcycles: int = 2
    n_pre_sweeps: int = 0
    n_post_sweeps: int = 2
    n_finest_sweeps: int = 2
    n_cells_coarsest: int = 10
    merge_levels: int = 1


@dataclass
class CaseConfig:
    start_time: float = 0.0
    end_time: float = 10.0
    delta_t: float = 1e-3
    write_interval: float = 0.05
    max_co: float = 0.5
    max_alpha_co: float = 0.5
    max_delta_t: float = 1.0
    adjust_time_step: bool = True
    write_binary: bool = True
    write_precision: int = 6
    time_precision: int = 6
    g: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    rho1: float = 998.2
    rho2: float = 1.0
    nu1: float = 1e-6
    nu2: float = 1.48e-5
    sigma: float = 0.0
    n_alpha_subcycles: int = 3
    n_alpha_corr: int = 1
    n_limiter_iter: int = 3
    c_alpha: float = 1.0
    n_correctors: int = 2
    n_non_orth: int = 0
    p_ref_point: np.ndarray = field(default_factory=lambda: np.zeros(3))
    p_ref_value: float = 0.0
    p_rgh: SolverControl = field(default_factory=SolverControl)
    p_rgh_final: SolverControl = field(default_factory=SolverControl)
    cofg: np.ndarray = field(default_factory=lambda: np.zeros(3))
    motion: np.ndarray | None = None  # (N,7) or None for a static mesh
    # per patch (mesh order)
    patch_bc_u: list = field(default_factory=list)
    patch_bc_alpha: list = field(default_factory=list)
    patch_bc_p: list = field(default_factory=list)
    patch_inlet_alpha: list = field(default_factory=list)
    patch_p0: list = field(default_factory=list)
    probes: np.ndarray | None = None  # (n,3) locations of the `probes` function object
    probe_fields: list = field(default_factory=list)
    n_subdomains: int = 1
    decomp_method: str = "scotch"


# ----------------------------------------------------------------------------------------
def _solver_control(d, name, path):
    sc = SolverControl()
    solver = lookup(d, "solver", f"{path}:{name}")
    sc.tolerance = to_float(lookup(d, "tolerance", f"{path}:{name}"))
    sc.rel_tol = to_float(lookup(d, "relTol", f"{path}:{name}", 0.0))
    sc.max_iter = int(lookup(d, "maxIter", path, 1000))

    def gamg_opts(g):
        sm = lookup(g, "smoother", f"{path}:{name}")
        if sm not in SMOOTHERS:
            raise FoamError(f"{path}:{name}: smoother '{sm}' is not supported ({', '.join(SMOOTHERS)})")
        sc.smoother = SMOOTHERS[sm]
        sc.n_vcycles = int(lookup(g, "nVcycles", path, 2))
        sc.n_pre_sweeps = int(lookup(g, "nPreSweeps", path, 0))
        sc.n_post_sweeps = int(lookup(g, "nPostSweeps", path, 2))
        sc.n_finest_sweeps = int(lookup(g, "nFinestSweeps", path, 2))
        sc.n_cells_coarsest = int(lookup(g, "nCellsInCoarsestLevel", path, 10))
        sc.merge_levels = int(lookup(g, "mergeLevels", path, 1))
        agg = lookup(g, "agglomerator", path, "faceAreaPair")
        if agg != "faceAreaPair":
            raise FoamError(f"{path}:{name}: agglomerator '{agg}' is not supported (faceAreaPair)")

    if solver == "GAMG":
        sc.type = 1
        gamg_opts(d)
    elif solver == "PCG":
        sc.type = 0
        pre = lookup(d, "preconditioner", f"{path}:{name}")
        if isinstance(pre, dict):
            kind = lookup(pre, "preconditioner", f"{path}:{name}.preconditioner")
            if kind == "GAMG":
                sc.precond = 1
                gamg_opts(pre)
            elif kind == "DIC":
                sc.precond = 0
            else:
                raise FoamError(f"{path}:{name}: preconditioner '{kind}' is not supported (GAMG, DIC)")
        elif pre == "DIC":
            sc.precond = 0
        else:
            raise FoamError(f"{path}:{name}: preconditioner '{pre}' is not supported (GAMG, DIC)")
    else:
        raise FoamError(f"{path}:{name}: solver '{solver}' is not supported (PCG, GAMG)")
    return sc


def _expect(val, allowed, what):
    v = " ".join(val) if isinstance(val, list) else val
    if v not in allowed:
        raise FoamError(f"{what}: '{v}' is not supported by this solver (supported: {' | '.join(allowed)})")
    return v


def read_config(case_dir, mesh=None):
    """Parse system/ and constant/ dictionaries (everything but mesh and fields)."""
    cfg = CaseConfig()
    p = os.path.join(case_dir, "system", "controlDict")
    cd = ff.read_dict(p)
    _expect(lookup(cd, "solver", p), ["incompressibleVoF"], f"{p}:solver")
    cfg.end_time = to_float(lookup(cd, "endTime", p))
    cfg.delta_t = to_float(lookup(cd, "deltaT", p))
    wc = lookup(cd, "writeControl", p)
    _expect(wc, ["adjustableRunTime"], f"{p}:writeControl")
    cfg.write_interval = to_float(lookup(cd, "writeInterval", p))
    cfg.adjust_time_step = to_bool(lookup(cd, "adjustTimeStep", p, "no"))
    cfg.max_co = to_float(lookup(cd, "maxCo", p, 1.0))
    cfg.max_alpha_co = to_float(lookup(cd, "maxAlphaCo", p, 1.0))
    cfg.max_delta_t = to_float(lookup(cd, "maxDeltaT", p, 1e30))
    cfg.write_binary = lookup(cd, "writeFormat", p, "ascii") == "binary"
    cfg.write_precision = int(lookup(cd, "writePrecision", p, 6))
    cfg.time_precision = int(lookup(cd, "timePrecision", p, 6))
    _expect(lookup(cd, "timeFormat", p, "general"), ["general"], f"{p}:timeFormat")
    start_from = lookup(cd, "startFrom", p, "latestTime")
    _expect(start_from, ["latestTime", "startTime"], f"{p}:startFrom")
    cfg.start_from = start_from
    cfg.start_time_entry = to_float(lookup(cd, "startTime", p, 0.0))

    p = os.path.join(case_dir, "system", "fvSchemes")
    fs = ff.read_dict(p)
    _expect(lookup(lookup(fs, "ddtSchemes", p), "default", p), ["Euler"], f"{p}:ddtSchemes")
    _expect(lookup(lookup(fs, "gradSchemes", p), "default", p), ["Gauss linear"], f"{p}:gradSchemes")
    div = lookup(fs, "divSchemes", p)
    _expect(lookup(div, "div(rhoPhi,U)", p), ["Gauss vanLeerV"], f"{p}:div(rhoPhi,U)")
    da = lookup(div, "div(phi,alpha)", p)
    if not (isinstance(da, list) and da[:3] == ["Gauss", "interfaceCompression", "vanLeer"] and len(da) == 4):
        raise FoamError(f"{p}:div(phi,alpha): only 'Gauss interfaceCompression vanLeer <cAlpha>' is supported, got {da!r}")
    cfg.c_alpha = float(da[3])
    _expect(lookup(div, "div(((rho*nuEff)*dev2(T(grad(U)))))", p), ["Gauss linear"], f"{p}:div(((rho*nuEff)*dev2(T(grad(U)))))")
    _expect(lookup(lookup(fs, "laplacianSchemes", p), "default", p), ["Gauss linear corrected"], f"{p}:laplacianSchemes")
    _expect(lookup(lookup(fs, "interpolationSchemes", p), "default", p), ["linear"], f"{p}:interpolationSchemes")
    _expect(lookup(lookup(fs, "snGradSchemes", p), "default", p), ["corrected"], f"{p}:snGradSchemes")

    p = os.path.join(case_dir, "system", "fvSolution")
    fv = ff.read_dict(p)
    sol = lookup(fv, "solvers", p)
    a = lookup(sol, "alpha.water", p)
    cfg.n_alpha_subcycles = int(lookup(a, "nSubCycles", p, lookup(a, "nAlphaSubCycles", p, 1)))
    cfg.n_alpha_corr = int(lookup(a, "nCorrectors", p, lookup(a, "nAlphaCorr", p, 1)))
    cfg.n_limiter_iter = int(lookup(a, "nLimiterIter", p, 3))
    if to_bool(lookup(a, "MULESCorr", p, "no")):
        raise FoamError(f"{p}:alpha.water: MULESCorr yes (semi-implicit MULES) is not supported")
    cfg.p_rgh = _solver_control(lookup(sol, "p_rgh", p), "p_rgh", p)
    cfg.p_rgh_final = _solver_control(lookup(sol, "p_rghFinal", p), "p_rghFinal", p)
    pim = lookup(fv, "PIMPLE", p)
    if to_bool(lookup(pim, "momentumPredictor", p, "yes")):
        raise FoamError(f"{p}:PIMPLE: momentumPredictor yes is not supported (the reference runs with 'no')")
    if int(lookup(pim, "nOuterCorrectors", p, 1)) != 1:
        raise FoamError(f"{p}:PIMPLE: nOuterCorrectors != 1 is not supported")
    if to_bool(lookup(pim, "correctPhi", p, "yes")):
        raise FoamError(f"{p}:PIMPLE: correctPhi yes is not supported (the reference runs with 'no')")
    cfg.n_correctors = int(lookup(pim, "nCorrectors", p, 1))
    cfg.n_non_orth = int(lookup(pim, "nNonOrthogonalCorrectors", p, 0))
    if "pRefPoint" in pim:
        cfg.p_ref_point = to_vector(pim["pRefPoint"], f"{p}:pRefPoint")
        cfg.p_ref_value = to_float(lookup(pim, "pRefValue", p))

    p = os.path.join(case_dir, "constant", "g")
    cfg.g = to_vector(lookup(ff.read_dict(p), "value", p), f"{p}:value")
    p = os.path.join(case_dir, "constant", "momentumTransport")
    _expect(lookup(ff.read_dict(p), "simulationType", p), ["laminar"], f"{p}:simulationType")
    p = os.path.join(case_dir, "constant", "phaseProperties")
    pp = ff.read_dict(p)
    phases = lookup(pp, "phases", p)
    if phases != ["water", "air"]:
        raise FoamError(f"{p}:phases: expected (water air), got {phases!r}")
    cfg.sigma = to_float(lookup(pp, "sigma", p))
    # the reference runs sigma 0 everywhere (constant/phaseProperties:19); sigma > 0 switches on the continuum
    # surface force with the walls' zeroGradient alpha (0/alpha.water:22-25), i.e. no contact-angle model
    if not (cfg.sigma >= 0.0 and np.isfinite(cfg.sigma)):
        raise FoamError(f"{p}:sigma: expected a non-negative surface tension coefficient, got {cfg.sigma}")
    for ph, rk, nk in (("water", "rho1", "nu1"), ("air", "rho2", "nu2")):
        p = os.path.join(case_dir, "constant", f"physicalProperties.{ph}")
        d = ff.read_dict(p)
        _expect(lookup(d, "viscosityModel", p), ["constant"], f"{p}:viscosityModel")
        setattr(cfg, rk, to_float(lookup(d, "rho", p)))
        setattr(cfg, nk, to_float(lookup(d, "nu", p)))

    p = os.path.join(case_dir, "constant", "dynamicMeshDict")
    if os.path.exists(p):
        mv = lookup(ff.read_dict(p), "mover", p)
        _expect(lookup(mv, "motionSolver", p), ["solidBody"], f"{p}:motionSolver")
        _expect(lookup(mv, "solidBodyMotionFunction", p), ["sixDoFMotion"], f"{p}:solidBodyMotionFunction")
        cfg.cofg = to_vector(lookup(mv, "CofG", p), f"{p}:CofG")
        files = set()
        for key, cols in (("translation", ["0", "1"]), ("rotation", ["0", "2"])):
            e = lookup(mv, key, p)
            _expect(lookup(e, "type", p), ["table"], f"{p}:{key}.type")
            if lookup(e, "columns", p) != cols:
                raise FoamError(f"{p}:{key}.columns: expected ({' '.join(cols)})")
            files.add(lookup(e, "file", p).strip('"').replace("$FOAM_CASE", case_dir))
        if len(files) != 1:
            raise FoamError(f"{p}: translation and rotation must read the same table file")
        cfg.motion = motion.read_table(files.pop())
        zone = lookup(mv, "cellZone", p)
        if mesh is not None and mesh.cell_zones:
            if zone not in mesh.cell_zones:
                raise FoamError(f"{p}:cellZone '{zone}' is not in constant/polyMesh/cellZones")
            if mesh.cell_zones[zone].size != mesh.n_cells:
                raise FoamError(f"{p}:cellZone '{zone}' does not cover the whole mesh; partial solid-body zones are not supported")

    p = os.path.join(case_dir, "system", "functions")
    if os.path.exists(p):
        for name, fo in ff.read_dict(p).items():
            if name == "FoamFile" or not isinstance(fo, dict):
                continue
            typ = lookup(fo, "type", p)
            if typ != "probes":
                raise FoamError(f"{p}:{name}: function object type '{typ}' is not supported (probes)")
            cfg.probes = np.array([[float(c) for c in v] for v in lookup(fo, "probeLocations", p)])
            cfg.probe_fields = list(lookup(fo, "fields", p))
            # the solver samples p (system/functions:28-31 of the reference); any other selection
            # would be written under the wrong name, so it is an error, not a silent default
            if cfg.probe_fields != ["p"]:
                raise FoamError(f"{p}:{name}: probes 'fields ({' '.join(map(str, cfg.probe_fields))})' is not supported: exactly 'fields (p)'")
    p = os.path.join(case_dir, "system", "decomposeParDict")
    if os.path.exists(p):
        dp = ff.read_dict(p)
        cfg.n_subdomains = int(lookup(dp, "numberOfSubdomains", p, 1))
        cfg.decomp_method = lookup(dp, "method", p, "scotch")
    return cfg


def _bc_tables(cfg, mesh, fields, tdir):
    cfg.patch_bc_u, cfg.patch_bc_alpha, cfg.patch_bc_p = [], [], []
    cfg.patch_inlet_alpha, cfg.patch_p0 = [], []
    for pt in mesh.patches:
        nm = pt["name"]
        if pt["type"] == "processor":
            cfg.patch_bc_u.append(-1)
            cfg.patch_bc_alpha.append(-1)
            cfg.patch_bc_p.append(-1)
            cfg.patch_inlet_alpha.append(0.0)
            cfg.patch_p0.append(0.0)
            continue
        for fld, table, out in (("U", U_BC, cfg.patch_bc_u), ("alpha.water", A_BC, cfg.patch_bc_alpha), ("p_rgh", P_BC, cfg.patch_bc_p)):
            bf = fields[fld].boundary
            e = lookup(bf, nm, f"{tdir}/{fld}:boundaryField")
            t = lookup(e, "type", f"{tdir}/{fld}:{nm}")
            if t == "noSlip" and fld == "U":
                t = "movingWallVelocity"  # identical on a wall that moves with the mesh
            if t not in table:
                raise FoamError(f"{tdir}/{fld}: boundary condition '{t}' on patch '{nm}' is not supported ({', '.join(table)})")
            out.append(table[t])
        e = fields["alpha.water"].boundary[nm]
        cfg.patch_inlet_alpha.append(to_float(e["inletValue"]) if "inletValue" in e else 0.0)
        e = fields["p_rgh"].boundary[nm]
        cfg.patch_p0.append(to_float(e["p0"]) if "p0" in e else 0.0)


def latest_time(case_dir, need="alpha.water"):
    """Newest COMPLETE time directory.  A directory the solver wrote is complete once
    `uniform/time` exists (write_time writes it last; a run killed mid-write leaves a directory
    without it, which `make resume` must not pick).  A hand-made initial directory (0/ after
    setFields: no `phi`) has no uniform/ and counts as soon as it holds `need`."""
    best = None
    for v, nm in ff.time_dirs(case_dir):
        d = os.path.join(case_dir, nm)
        if not os.path.exists(os.path.join(d, need)):
            continue
        if v > 0 and os.path.exists(os.path.join(d, "phi")) and not os.path.exists(os.path.join(d, "uniform", "time")):
            continue  # partially written
        best = (v, nm)
    if best is None:
        raise FoamError(f"{case_dir}: no time directory holds {need}")
    return best


class Case:
    """A case directory loaded into memory: mesh, config, start fields."""

    def __init__(self, case_dir, processor=None):
        """processor = k: the rank's share of a decomposed case - mesh and fields from
        `processor<k>/` (as `foamRun -parallel` reads them), dictionaries from the case root;
        time directories are then written under `processor<k>/` too."""
        self.root = case_dir
        self.processor = processor
        if processor is not None:
            case_dir = os.path.join(case_dir, f"processor{processor}")
            if not os.path.isdir(case_dir):
                raise FoamError(f"{case_dir}: not found (run decomposePar first)")
        self.dir = case_dir
        self.mesh = ff.read_polymesh(case_dir)
        self.cfg = read_config(self.root, self.mesh)
        if self.cfg.start_from == "latestTime":
            self.start_value, self.start_name = latest_time(case_dir)
        else:
            self.start_value = self.cfg.start_time_entry
            self.start_name = ff.time_name(self.start_value, self.cfg.time_precision)
        self.cfg.start_time = self.start_value
        tdir = os.path.join(case_dir, self.start_name)
        self.fields = {}
        for nm in ("alpha.water", "U", "p_rgh"):
            self.fields[nm] = ff.read_field(os.path.join(tdir, nm))
        for nm in ("phi", "Uf", "p", "rho"):
            fp = os.path.join(tdir, nm)
            if os.path.exists(fp):
                self.fields[nm] = ff.read_field(fp)
        _bc_tables(self.cfg, self.mesh, self.fields, self.start_name)
        self.restart_delta_t = None
        up = os.path.join(tdir, "uniform", "time")
        if os.path.exists(up):
            self.restart_delta_t = to_float(lookup(ff.read_dict(up), "deltaT", up))


# ----------------------------------------------------------------------------------------
# case writers (same values as /root/reference/circularSloshingTank/{0,constant,system})
# ----------------------------------------------------------------------------------------
def _dict_file(path, cls, obj, location, body):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(ff._hdr(cls, obj, location))
        f.write(body)
        f.write(ff.END)


def write_template(case_dir, end_time=10.0, delta_t=0.001, write_interval=0.05, wall="walls", atmosphere="atmosphere", cell_zone="internalMesh", start_from="latestTime", n_subdomains=1, method="scotch", fill_z=0.5, fill_z_lo=-1.0, p_final_max_iter=20):
    """system/, constant/ and 0/ dictionaries with the reference's numerics
    (circularSloshingTank/system/controlDict:17-51, fvSchemes:17-48, fvSolution:17-95,
    constant/*, 0/*).  atmosphere=None writes the closed-tank variant of sloshingTank3D6DoF."""
    s, c, z = (os.path.join(case_dir, d) for d in ("system", "constant", "0"))
    _dict_file(
        os.path.join(s, "controlDict"), "dictionary", "controlDict", "system",
        f"solver          incompressibleVoF;\n\nstartFrom       {start_from};\n\nstartTime       0;\n\nstopAt          endTime;\n\n"
        f"endTime         {end_time:g};\n\ndeltaT          {delta_t:g};\n\nwriteControl    adjustableRunTime;\n\nwriteInterval   {write_interval:g};\n\n"
        "purgeWrite      0;\n\nwriteFormat     binary;\n\nwritePrecision  6;\n\nwriteCompression off;\n\ntimeFormat      general;\n\n"
        "timePrecision   6;\n\nrunTimeModifiable yes;\n\nadjustTimeStep  yes;\n\nmaxCo           0.5;\nmaxAlphaCo      0.5;\nmaxDeltaT       1;\n",
    )
    _dict_file(
        os.path.join(s, "fvSchemes"), "dictionary", "fvSchemes", "system",
        "ddtSchemes\n{\n    default         Euler;\n}\n\ngradSchemes\n{\n    default         Gauss linear;\n}\n\n"
        "divSchemes\n{\n    div(rhoPhi,U)  Gauss vanLeerV;\n    div(phi,alpha)  Gauss interfaceCompression vanLeer 1;\n"
        "    div(((rho*nuEff)*dev2(T(grad(U))))) Gauss linear;\n}\n\nlaplacianSchemes\n{\n    default         Gauss linear corrected;\n}\n\n"
        "interpolationSchemes\n{\n    default         linear;\n}\n\nsnGradSchemes\n{\n    default         corrected;\n}\n",
    )
    _dict_file(
        os.path.join(s, "fvSolution"), "dictionary", "fvSolution", "system",
        "solvers\n{\n    alpha.water\n    {\n        nCorrectors     1;\n        nSubCycles      3;\n    }\n\n"
        "    \"pcorr.*\"\n    {\n        solver          PCG;\n        preconditioner\n        {\n            preconditioner  GAMG;\n"
        "            tolerance       1e-05;\n            relTol          0;\n            smoother        DICGaussSeidel;\n            cacheAgglomeration no;\n        }\n"
        "        tolerance       1e-05;\n        relTol          0;\n        maxIter         100;\n    }\n\n"
        "    p_rgh\n    {\n        solver          GAMG;\n        tolerance       1e-08;\n        relTol          0.01;\n        smoother        DIC;\n    }\n\n"
        "    p_rghFinal\n    {\n        solver          PCG;\n        preconditioner\n        {\n            preconditioner  GAMG;\n            tolerance       2e-09;\n"
        "            relTol          0;\n            nVcycles        2;\n            smoother        DICGaussSeidel;\n            nPreSweeps      2;\n        }\n"
        f"        tolerance       2e-09;\n        relTol          0;\n        maxIter         {p_final_max_iter};\n    }}\n\n"
        "    U\n    {\n        solver          smoothSolver;\n        smoother        GaussSeidel;\n        tolerance       1e-06;\n        relTol          0;\n        nSweeps         1;\n    }\n}\n\n"
        "PIMPLE\n{\n    momentumPredictor no;\n    nCorrectors     2;\n    nNonOrthogonalCorrectors 0;\n    correctPhi      no;\n\n"
        "    pRefPoint       (0 0 0.15);\n    pRefValue       1e5;\n}\n\nrelaxationFactors\n{\n    equations\n    {\n        \".*\"            1;\n    }\n}\n",
    )
    _dict_file(
        os.path.join(s, "functions"), "dictionary", "functions", "system",
        "probes\n{\n    type            probes;\n    libs            (\"libsampling.so\");\n    writeControl   timeStep;\n    writeInterval  1;\n"
        "    probeLocations\n    (\n        (0 9.95 19.77)\n        (0 -9.95 19.77)\n    );\n    fixedLocations  false;\n    fields\n    (\n        p\n    );\n}\n",
    )
    _dict_file(
        os.path.join(s, "decomposeParDict"), "dictionary", "decomposeParDict", "system",
        f"numberOfSubdomains {n_subdomains};\n\nmethod          {method};\n\nsimpleCoeffs\n{{\n    n               (1 1 {n_subdomains});\n}}\n",
    )
    write_setfields_dict(case_dir, fill_z, fill_z_lo)
    _dict_file(
        os.path.join(c, "dynamicMeshDict"), "dictionary", "dynamicMeshDict", "constant",
        f"mover\n{{\n    type            motionSolver;\n\n    libs            (\"libfvMotionSolvers.so\");\n\n    motionSolver    solidBody;\n\n    cellZone        {cell_zone};\n\n"
        "    solidBodyMotionFunction sixDoFMotion;\n\n    CofG            (0 0 0);\n\n    translation\n    {\n        type            table;\n"
        "        file            \"$FOAM_CASE/constant/6DoF.dat\";\n        columns         (0 1);\n    }\n\n    rotation\n    {\n        type            table;\n"
        "        file            \"$FOAM_CASE/constant/6DoF.dat\";\n        columns         (0 2);\n    }\n}\n",
    )
    _dict_file(os.path.join(c, "g"), "uniformDimensionedVectorField", "g", "constant", "dimensions      [0 1 -2 0 0 0 0];\nvalue           (0 0 -9.81);\n")
    _dict_file(os.path.join(c, "momentumTransport"), "dictionary", "momentumTransport", "constant", "simulationType  laminar;\n")
    _dict_file(os.path.join(c, "phaseProperties"), "dictionary", "phaseProperties", "constant", "phases          (water air);\n\nsigma           0;\n")
    _dict_file(os.path.join(c, "physicalProperties.water"), "dictionary", "physicalProperties.water", "constant", "viscosityModel  constant;\n\nnu              1e-06;\n\nrho             998.2;\n")
    _dict_file(os.path.join(c, "physicalProperties.air"), "dictionary", "physicalProperties.air", "constant", "viscosityModel  constant;\n\nnu              1.48e-05;\n\nrho             1;\n")

    def patches(wall_body, atm_body):
        t = f"    {wall}\n    {{\n{wall_body}    }}\n"
        if atmosphere:
            t += f"\n    {atmosphere}\n    {{\n{atm_body}    }}\n"
        return t

    _dict_file(
        os.path.join(z, "U"), "volVectorField", "U", None,
        "dimensions      [0 1 -1 0 0 0 0];\n\ninternalField   uniform (0 0 0);\n\nboundaryField\n{\n"
        + patches("        type            movingWallVelocity;\n        value           uniform (0 0 0);\n", "        type            pressureInletOutletVelocity;\n        value           uniform (0 0 0);\n")
        + "}\n",
    )
    _dict_file(
        os.path.join(z, "alpha.water"), "volScalarField", "alpha.water", None,
        "dimensions      [0 0 0 0 0 0 0];\n\ninternalField   uniform 0;\n\nboundaryField\n{\n"
        + patches("        type            zeroGradient;\n", "        type            inletOutlet;\n        inletValue      uniform 0;\n        value           uniform 0;\n")
        + "}\n",
    )
    _dict_file(
        os.path.join(z, "p_rgh"), "volScalarField", "p_rgh", None,
        "dimensions      [1 -1 -2 0 0 0 0];\n\ninternalField   uniform 0;\n\nboundaryField\n{\n"
        + patches("        type            fixedFluxPressure;\n        value           uniform 0;\n", "        type            totalPressure;\n        p0              uniform 0;\n")
        + "}\n",
    )


def write_setfields_dict(case_dir, fill_z, z_lo=-1.0):
    """system/setFieldsDict as update_setFields.py:21-37 writes it (box up to z = fill_z)."""
    _dict_file(
        os.path.join(case_dir, "system", "setFieldsDict"), "dictionary", "setFieldsDict", "system",
        "defaultValues\n{\n    alpha.water 0;\n}\n\nzones\n{\n    water\n    {\n        type box;\n"
        f"        box (-100 -100 {z_lo}) (100 100 {fill_z});\n\n        values\n        {{\n            alpha.water 1;\n        }}\n    }}\n}}\n",
    )


def set_fields(case_dir, time_name="0"):
    """Box initialisation of alpha.water by cell centre: the `setFields` step of
    circularSloshingTank/Makefile:74 for the dictionary update_setFields.py writes."""
    p = os.path.join(case_dir, "system", "setFieldsDict")
    d = ff.read_dict(p)
    default = to_float(lookup(lookup(d, "defaultValues", p), "alpha.water", p))
    mesh = ff.read_polymesh(case_dir)
    C, V = meshgen.cell_geometry(mesh)
    a = np.full(mesh.n_cells, default)
    for name, zn in lookup(d, "zones", p).items():
        _expect(lookup(zn, "type", p), ["box"], f"{p}:{name}.type")
        lo, hi = (np.array([float(c) for c in v]) for v in lookup(zn, "box", p))
        val = to_float(lookup(lookup(zn, "values", p), "alpha.water", p))
        inside = np.all((C >= lo) & (C <= hi), axis=1)
        a[inside] = val
    fp = os.path.join(case_dir, time_name, "alpha.water")
    fld = ff.read_field(fp)
    fld.internal = a
    for e in fld.boundary.values():
        e.pop("value", None)
        if e.get("type") == "inletOutlet":
            e["value"] = 0.0
    ff.write_field(fp, fld, binary=True, location=time_name)
    return a


def setup_case(case_dir, H=0.1, D=0.02, geo="flat", R=0.003, freq=2.0, duration=10.0, dt=0.001, ramp=-1, n_rings=8, n_layers=16, cell="tet", end_time=None, write_interval=0.05, overwrite=True, p_final_max_iter=20):
    """Stand-in for main.py:setup_case (main.py:266-331): template + generate_motion.py +
    update_setFields.py + mesh (n_rings/n_layers replace gmsh's lc) + setFields."""
    if overwrite and os.path.isdir(case_dir):
        shutil.rmtree(case_dir)
    write_template(case_dir, end_time=duration if end_time is None else end_time, write_interval=write_interval, fill_z=H / 2.0, p_final_max_iter=p_final_max_iter)
    rows = motion.orbital_table(R, freq, duration, dt, ramp)
    motion.write_table(os.path.join(case_dir, "constant", "6DoF.dat"), rows)
    mesh = meshgen.cylinder_mesh(H, D, n_rings, n_layers, geo, cell)
    ff.write_polymesh(case_dir, mesh, binary=True)
    set_fields(case_dir)
    with open(os.path.join(case_dir, "case.foam"), "w"):
        pass
    return case_dir


def setup_tutorial_case(case_dir, nx=10, ny=20, nz=15, end_time=40.0, overwrite=True, p_final_max_iter=20):
    """sloshingTank3D6DoF: closed tank, single patch `wall`, zone `all`, fill z < 0,
    startFrom startTime, deltaT 0.01, 6DoF table from the gen6DoF restatement
    (sloshingTank3D6DoF/system/controlDict:19-27, setFieldsDict:26-29, gen6DoF.C:44-82)."""
    if overwrite and os.path.isdir(case_dir):
        shutil.rmtree(case_dir)
    write_template(case_dir, end_time=end_time, delta_t=0.01, wall="wall", atmosphere=None, cell_zone="all", start_from="startTime", fill_z=0.0, fill_z_lo=-100.0, p_final_max_iter=p_final_max_iter)
    motion.write_table(os.path.join(case_dir, "constant", "6DoF.dat"), motion.gen6dof_table())
    ff.write_polymesh(case_dir, meshgen.sloshing_tank3d_mesh(nx, ny, nz), binary=True)
    set_fields(case_dir)
    return case_dir


def main(argv=None):
    """`setFields [-case DIR]` (circularSloshingTank/Makefile:74) for the box dictionary
    update_setFields.py writes."""
    import sys

    argv = list(sys.argv[1:] if argv is None else argv)
    if argv and argv[0] == "setFields":
        argv.pop(0)
    case_dir = os.getcwd()
    while argv:
        a = argv.pop(0)
        if a == "-case":
            case_dir = argv.pop(0)
        else:
            raise SystemExit(f"setFields (tppvof): unknown option {a}")
    try:
        set_fields(case_dir)
    except Exception as e:
        print(f"--> FOAM FATAL ERROR: {e}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    import sys

    sys.exit(main())
