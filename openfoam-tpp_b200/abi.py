"""ctypes mirrors of the C structs in include/tppvof.h (and of the oracle's identical
layout in oracle/vof_oracle.h), filled from a PolyMesh + CaseConfig."""
from __future__ import annotations

import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class MeshStruct(C.Structure):
    _fields_ = [
        ("n_points", C.c_int), ("n_faces", C.c_int), ("n_internal", C.c_int), ("n_cells", C.c_int), ("n_patches", C.c_int),
        ("points", c_double_p), ("face_offsets", c_int_p), ("face_labels", c_int_p), ("owner", c_int_p), ("neighbour", c_int_p),
        ("patch_start", c_int_p), ("patch_size", c_int_p), ("patch_bc_u", c_int_p), ("patch_bc_alpha", c_int_p), ("patch_bc_p", c_int_p),
        ("patch_inlet_alpha", c_double_p), ("patch_p0", c_double_p),
        ("patch_neighb_proc", c_int_p),  # tpp only (the oracle's struct is the prefix before this field)
    ]


class SolverStruct(C.Structure):
    _fields_ = [
        ("type", C.c_int), ("precond", C.c_int), ("smoother", C.c_int),
        ("tolerance", C.c_double), ("rel_tol", C.c_double), ("max_iter", C.c_int),
        ("n_vcycles", C.c_int), ("n_pre_sweeps", C.c_int), ("n_post_sweeps", C.c_int), ("n_finest_sweeps", C.c_int),
        ("n_cells_coarsest", C.c_int), ("merge_levels", C.c_int),
    ]


class ConfigStruct(C.Structure):
    _fields_ = [
        ("start_time", C.c_double), ("end_time", C.c_double), ("delta_t", C.c_double), ("write_interval", C.c_double),
        ("max_co", C.c_double), ("max_alpha_co", C.c_double), ("max_delta_t", C.c_double), ("adjust_time_step", C.c_int),
        ("g", C.c_double * 3), ("rho1", C.c_double), ("rho2", C.c_double), ("nu1", C.c_double), ("nu2", C.c_double), ("sigma", C.c_double),
        ("n_alpha_subcycles", C.c_int), ("n_alpha_corr", C.c_int), ("n_limiter_iter", C.c_int), ("c_alpha", C.c_double),
        ("n_correctors", C.c_int), ("n_non_orth", C.c_int), ("p_ref_point", C.c_double * 3), ("p_ref_value", C.c_double),
        ("p_rgh", SolverStruct), ("p_rgh_final", SolverStruct),
        ("cofg", C.c_double * 3), ("n_motion", C.c_int), ("motion", c_double_p),
    ]


def solver_struct(sc):
    return SolverStruct(sc.type, sc.precond, sc.smoother, sc.tolerance, sc.rel_tol, sc.max_iter, sc.n_vcycles, sc.n_pre_sweeps,
                        sc.n_post_sweeps, sc.n_finest_sweeps, sc.n_cells_coarsest, sc.merge_levels)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def build_structs(mesh, cfg):
    """Returns (MeshStruct, ConfigStruct, keepalive list of numpy arrays)."""
    keep = []

    def arr(x, dt):
        a = np.ascontiguousarray(x, dtype=dt)
        keep.append(a)
        return a

    # the C struct carries no length for face_labels: check the offsets against it before the library indexes with them
    off, nlab = np.asarray(mesh.face_offsets), len(mesh.face_labels)
    if off.size != mesh.n_faces + 1 or off[0] != 0 or off[-1] > nlab or np.any(np.diff(off) < 0):
        raise ValueError(f"polyMesh faces: the offsets do not span the {nlab} point labels")
    pts = arr(mesh.points, np.float64)
    m = MeshStruct()
    m.n_points, m.n_faces, m.n_internal, m.n_cells, m.n_patches = mesh.n_points, mesh.n_faces, mesh.n_internal, mesh.n_cells, len(mesh.patches)
    m.points = _dp(pts)
    m.face_offsets = _ip(arr(mesh.face_offsets, np.int32))
    m.face_labels = _ip(arr(mesh.face_labels, np.int32))
    m.owner = _ip(arr(mesh.owner, np.int32))
    m.neighbour = _ip(arr(mesh.neighbour, np.int32))
    m.patch_start = _ip(arr([p["startFace"] for p in mesh.patches], np.int32))
    m.patch_size = _ip(arr([p["nFaces"] for p in mesh.patches], np.int32))
    m.patch_bc_u = _ip(arr(cfg.patch_bc_u, np.int32))
    m.patch_bc_alpha = _ip(arr(cfg.patch_bc_alpha, np.int32))
    m.patch_bc_p = _ip(arr(cfg.patch_bc_p, np.int32))
    m.patch_inlet_alpha = _dp(arr(cfg.patch_inlet_alpha, np.float64))
    m.patch_p0 = _dp(arr(cfg.patch_p0, np.float64))
    m.patch_neighb_proc = _ip(arr([p.get("neighbProcNo", -1) for p in mesh.patches], np.int32))
    c = ConfigStruct()
    c.start_time, c.end_time, c.delta_t, c.write_interval = cfg.start_time, cfg.end_time, cfg.delta_t, cfg.write_interval
    c.max_co, c.max_alpha_co, c.max_delta_t, c.adjust_time_step = cfg.max_co, cfg.max_alpha_co, cfg.max_delta_t, int(cfg.adjust_time_step)
    c.g = (C.c_double * 3)(*cfg.g)
    c.rho1, c.rho2, c.nu1, c.nu2, c.sigma = cfg.rho1, cfg.rho2, cfg.nu1, cfg.nu2, cfg.sigma
    c.n_alpha_subcycles, c.n_alpha_corr, c.n_limiter_iter, c.c_alpha = cfg.n_alpha_subcycles, cfg.n_alpha_corr, cfg.n_limiter_iter, cfg.c_alpha
    c.n_correctors, c.n_non_orth = cfg.n_correctors, cfg.n_non_orth
    c.p_ref_point = (C.c_double * 3)(*cfg.p_ref_point)
    c.p_ref_value = cfg.p_ref_value
    c.p_rgh, c.p_rgh_final = solver_struct(cfg.p_rgh), solver_struct(cfg.p_rgh_final)
    c.cofg = (C.c_double * 3)(*cfg.cofg)
    if cfg.motion is not None and len(cfg.motion):
        mt = arr(cfg.motion, np.float64)
        c.n_motion, c.motion = mt.shape[0], _dp(mt)
    else:
        c.n_motion, c.motion = 0, None
    return m, c, keep
