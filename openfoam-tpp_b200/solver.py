"""ctypes binding of libtppvof.so (include/tppvof.h) — the thin host layer over the CUDA solver.

There is no CPU path: `Solver` raises when the CUDA library is missing or no GPU is usable.
(`lib_path` exists so the unit tests can point the same binding at tests/_emu's host
emulation of the kernel bodies; the package itself never does.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtppvof.so")
_LIBS = {}

INFO_KEYS = ["t", "dt", "step", "Co", "alphaCo", "it0", "r00", "r0", "it1", "r01", "r1", "refCell", "deltaN", "writeIndex", "levels", "launches"]


class SolverError(RuntimeError):
    pass


XCB = C.CFUNCTYPE(C.c_int, C.c_void_p, abi.c_double_p, abi.c_double_p, C.c_int)
RCB = C.CFUNCTYPE(C.c_int, C.c_void_p, abi.c_double_p, C.c_int, C.c_int)


def nccl_library_path():
    """The libnccl the process already has loaded (torch's bundled copy), from /proc/self/maps."""
    import torch  # noqa: F401  (loads it)

    with open("/proc/self/maps") as f:
        for line in f:
            if "libnccl" in line:
                return line.split()[-1]
    import glob

    hits = glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib", "libnccl.so*"))
    if hits:
        return hits[0]
    raise SolverError("NCCL library not found")


def load(lib_path=None):
    path = lib_path or LIB_PATH
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise SolverError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    H = C.c_void_p
    L.tpp_create.argtypes = [C.POINTER(abi.MeshStruct), C.POINTER(abi.ConfigStruct), C.c_int, C.POINTER(H)]
    L.tpp_destroy.argtypes = [H]
    L.tpp_last_error.restype = C.c_char_p
    L.tpp_version.restype = C.c_char_p
    L.tpp_size.restype = C.c_long
    L.tpp_size.argtypes = [H, C.c_char_p]
    for f in (L.tpp_get, L.tpp_set):
        f.restype = C.c_long
        f.argtypes = [H, C.c_char_p, abi.c_double_p, C.c_long]
    L.tpp_device_ptr.argtypes = [H, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_long)]
    L.tpp_init_fields.argtypes = [H]
    L.tpp_set_delta_t.argtypes = [H, C.c_double]
    L.tpp_set_time.argtypes = [H, C.c_double, C.c_double]
    L.tpp_step.argtypes = [H, C.c_int]
    L.tpp_run_to_write.argtypes = [H, C.c_long]
    L.tpp_stats.argtypes = [H, C.c_int, abi.c_double_p]
    L.tpp_interface.argtypes = [H, C.c_double, abi.c_double_p]
    L.tpp_get_async.restype = C.c_long
    L.tpp_get_async.argtypes = [H, C.c_char_p, abi.c_double_p, C.c_long]
    L.tpp_sync.argtypes = [H]
    L.tpp_get_int.restype = C.c_long
    L.tpp_get_int.argtypes = [H, C.c_char_p, abi.c_int_p, C.c_long]
    L.tpp_stage.argtypes = [H, C.c_char_p]
    L.tpp_info.argtypes = [H, abi.c_double_p]
    L.tpp_solve.argtypes = [H, C.POINTER(abi.SolverStruct)] + [abi.c_double_p] * 6
    L.tpp_set_probes.argtypes = [H, C.c_int, abi.c_int_p]
    L.tpp_probe_log.restype = C.c_long
    L.tpp_probe_log.argtypes = [H, abi.c_double_p, C.c_long]
    L.tpp_find_cell.argtypes = [H, abi.c_double_p]
    L.tpp_use_stream.argtypes = [H, C.c_void_p]
    L.tpp_profile.argtypes = [H, C.c_int]
    L.tpp_profile_report.restype = C.c_long
    L.tpp_profile_report.argtypes = [H, C.c_char_p, C.c_long]
    L.tpp_nccl_unique_id.argtypes = [C.c_char_p, C.c_char_p]
    L.tpp_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
    L.tpp_comm_callbacks.argtypes = [H, C.c_int, C.c_int, XCB, RCB, C.c_void_p]
    L.tpp_amg_levels.argtypes = [H, abi.c_int_p, abi.c_int_p, C.c_int]
    L.tpp_amg_layout.argtypes = [H, abi.c_int_p]
    L.tpp_ghost_layout.argtypes = [H, abi.c_int_p, abi.c_int_p, abi.c_int_p, abi.c_int_p, abi.c_int_p, C.c_int]
    L.tpp_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(H)]
    L.tpp_case_start.argtypes = [H]
    L.tpp_write_time.argtypes = [H]
    L.tpp_run_case.restype = C.c_long
    L.tpp_run_case.argtypes = [H, C.c_long, C.c_int]
    L.tpp_case_query.restype = C.c_long
    L.tpp_case_query.argtypes = [H, C.c_char_p, C.c_char_p, C.c_long]
    L.tpp_read_field.restype = C.c_long
    L.tpp_read_field.argtypes = [C.c_char_p, abi.c_double_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _LIBS[path] = L
    return L


def read_field_file(path, lib_path=None):
    """internalField of a field file through the library's reader (tpp_read_field): (array, uniform)."""
    L = load(lib_path)
    nc, uni = C.c_int(), C.c_int()
    # one pass: no field holds more doubles than half its file's bytes (untouched pages of the buffer cost nothing)
    cap = (os.path.getsize(path) if os.path.exists(path) else 0) // 2 + 16
    a = np.empty(cap)
    n = L.tpp_read_field(os.fsencode(path), a.ctypes.data_as(abi.c_double_p), cap, C.byref(nc), C.byref(uni))
    if n < 0:
        raise SolverError(f"tpp_read_field failed ({n}): {L.tpp_last_error().decode()}")
    assert n <= cap
    a = a[:n].copy()
    return (a.reshape(-1, nc.value) if nc.value > 1 else a), bool(uni.value)


class Solver:
    """One case on one GPU."""

    def __init__(self, mesh, cfg, device=0, lib_path=None):
        self.L = load(lib_path)
        try:
            m, c, self._keep = abi.build_structs(mesh, cfg)
        except ValueError as e:  # what the array ABI cannot check for itself (no length for face_labels)
            raise SolverError(f"invalid mesh / configuration: {e}") from None
        h = C.c_void_p()
        rc = self.L.tpp_create(C.byref(m), C.byref(c), device, C.byref(h))
        if rc != 0:
            raise SolverError(f"tpp_create failed ({rc}): {self.L.tpp_last_error().decode()}")
        self.h = h
        self.mesh, self.cfg = mesh, cfg
        self._nprobe = 0

    @classmethod
    def open(cls, case_dir, device=0, lib_path=None, processor=-1):
        """The library's own case reader (tpp_open, include/tppvof.h): what a host without a FoamFile
        parser binds.  The Python host normally reads the case itself (case.Case) and uses __init__; the
        two must give the same solver state (tests/test_caseio.py)."""
        self = cls.__new__(cls)
        self.L = load(lib_path)
        self._keep, self._nprobe = [], 0
        h = C.c_void_p()
        rc = self.L.tpp_open(os.fsencode(case_dir), processor, device, C.byref(h))
        if rc != 0:
            raise SolverError(f"tpp_open failed ({rc}): {self.L.tpp_last_error().decode()}")
        self.h = h
        self.mesh = self.cfg = None
        return self

    def case_query(self, what):
        buf = C.create_string_buffer(4096)
        n = self.L.tpp_case_query(self.h, what.encode(), buf, 4096)
        if n < 0:
            self._err("tpp_case_query")
        return buf.value.decode() if what in ("start_time", "time", "dir") else int(n)

    def case_start(self):
        if self.L.tpp_case_start(self.h) != 0:
            self._err("tpp_case_start")

    def write_time(self):
        if self.L.tpp_write_time(self.h) != 0:
            self._err("tpp_write_time")

    def run_case(self, max_steps=-1, verbose=False, interface=False):
        n = self.L.tpp_run_case(self.h, max_steps, int(bool(verbose)) | (2 if interface else 0))
        if n < 0:
            self._err("tpp_run_case")
        return int(n)

    def close(self):
        if getattr(self, "h", None):
            self.L.tpp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, what):
        raise SolverError(f"{what}: {self.L.tpp_last_error().decode()}")

    def size(self, name):
        return self.L.tpp_size(self.h, name.encode())

    def get(self, name):
        n = self.size(name)
        if n < 0:
            raise KeyError(name)
        a = np.empty(n, dtype=np.float64)
        if self.L.tpp_get(self.h, name.encode(), a.ctypes.data_as(abi.c_double_p), n) < 0:
            self._err("tpp_get")
        return a

    def set(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        if self.L.tpp_set(self.h, name.encode(), a.ctypes.data_as(abi.c_double_p), a.size) < 0:
            self._err("tpp_set")

    def device_ptr(self, name):
        p, n = C.c_void_p(), C.c_long()
        if self.L.tpp_device_ptr(self.h, name.encode(), C.byref(p), C.byref(n)) != 0:
            self._err("tpp_device_ptr")
        return p.value, n.value

    def init_fields(self):
        self.L.tpp_init_fields(self.h)

    def set_delta_t(self, dt):
        self.L.tpp_set_delta_t(self.h, dt)

    def set_time(self, t, dt):
        self.L.tpp_set_time(self.h, t, dt)

    def stage(self, name):
        if self.L.tpp_stage(self.h, name.encode()) != 0:
            self._err("tpp_stage")

    def step(self, n=1):
        if self.L.tpp_step(self.h, n) != 0:
            self._err("tpp_step")

    def run_to_write(self, max_steps=10**9):
        """1: stopped at a write time, 0: endTime reached, 2: max_steps used up.  A failed run (lost
        halo, grid-barrier timeout, non-finite Courant number / residual, CUDA error) raises: foamRun
        exits non-zero, as OpenFOAM does and `check=True` expects (main.py:345)."""
        rc = self.L.tpp_run_to_write(self.h, max_steps)
        if rc < 0:
            self._err("tpp_run_to_write")
        return rc

    def info(self):
        o = np.zeros(16)
        self.L.tpp_info(self.h, o.ctypes.data_as(abi.c_double_p))
        return dict(zip(INFO_KEYS, o))

    def get_int(self, name):
        n = self.L.tpp_get_int(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        a = np.empty(n, dtype=np.int32)
        if self.L.tpp_get_int(self.h, name.encode(), a.ctypes.data_as(abi.c_int_p), n) < 0:
            self._err("tpp_get_int")
        return a

    def interface_summary(self, iso=0.5):
        """(time, max_z, min_z, mean_z, num_points) of the alpha = iso contour, computed on the device:
        the columns of the reference's interface_summary.csv (main.py:751,780)."""
        o = np.zeros(5)
        if self.L.tpp_interface(self.h, float(iso), o.ctypes.data_as(abi.c_double_p)) != 0:
            self._err("tpp_interface")
        return float(o[4]), float(o[0]), float(o[1]), float(o[2]), int(o[3])

    def stats(self, reset=-1):
        """tpp_stats: iteration statistics and alpha-volume balance since the last reset."""
        o = np.zeros(9)
        if self.L.tpp_stats(self.h, int(reset), o.ctypes.data_as(abi.c_double_p)) != 0:
            self._err("tpp_stats")
        n = max(o[0], 1.0)
        return {"steps": int(o[0]), "it0_mean": o[1] / n, "it1_mean": o[2] / n, "it0_max": int(o[3]), "it1_max": int(o[4]), "cap_hits": int(o[5]),
                "alpha_volume_start": o[6], "alpha_volume": o[7], "alpha_boundary_outflow": o[8],
                "alpha_balance_rel": abs(o[7] - o[6] + o[8]) / max(abs(o[6]), 1e-300)}

    def solve(self, ctl, diag, upper, b, x0=None):
        x = np.zeros(self.mesh.n_cells) if x0 is None else np.array(x0, dtype=np.float64)
        r0, r = C.c_double(), C.c_double()
        s = abi.solver_struct(ctl)
        d_, u_, b_ = (np.ascontiguousarray(v, dtype=np.float64) for v in (diag, upper, b))
        it = self.L.tpp_solve(self.h, C.byref(s), d_.ctypes.data_as(abi.c_double_p), u_.ctypes.data_as(abi.c_double_p), b_.ctypes.data_as(abi.c_double_p),
                              x.ctypes.data_as(abi.c_double_p), C.cast(C.byref(r0), abi.c_double_p), C.cast(C.byref(r), abi.c_double_p))
        return x, it, r0.value, r.value

    def set_probes(self, cells):
        a = np.ascontiguousarray(cells, dtype=np.int32)
        self._nprobe = a.size
        self.L.tpp_set_probes(self.h, a.size, a.ctypes.data_as(abi.c_int_p))

    def probe_log(self, cap=1 << 16):
        w = 1 + self._nprobe
        a = np.empty((cap, w))
        n = self.L.tpp_probe_log(self.h, a.ctypes.data_as(abi.c_double_p), cap)
        return a[:n].copy()

    def find_cell(self, xyz):
        a = np.ascontiguousarray(xyz, dtype=np.float64)
        return self.L.tpp_find_cell(self.h, a.ctypes.data_as(abi.c_double_p))

    def amg_levels(self):
        """[(rows, faces)] of the fine level and every coarse level of the cached hierarchy."""
        n, f = np.zeros(32, dtype=np.int32), np.zeros(32, dtype=np.int32)
        k = self.L.tpp_amg_levels(self.h, n.ctypes.data_as(abi.c_int_p), f.ctypes.data_as(abi.c_int_p), 32)
        return [(int(n[i]), int(f[i])) for i in range(k)]

    def amg_layout(self):
        """levels smoothed kernel by kernel (rows distributed), tail levels (gathered, one persistent
        kernel), rows of the first tail level, CTAs of the tail kernel"""
        o = np.zeros(4, dtype=np.int32)
        self.L.tpp_amg_layout(self.h, o.ctypes.data_as(abi.c_int_p))
        return dict(zip(("kernel_levels", "tail_levels", "tail_rows", "tail_ctas"), (int(x) for x in o)))

    def use_stream(self, cuda_stream):
        self.L.tpp_use_stream(self.h, C.c_void_p(cuda_stream))

    def profile(self, on=True):
        self.L.tpp_profile(self.h, int(on))

    def profile_report(self):
        """{kernel: (launches, total_ms)} since profiling was switched on / last report."""
        buf = C.create_string_buffer(1 << 16)
        n = self.L.tpp_profile_report(self.h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            k, c, ms = line.split()
            out[k] = (int(c), float(ms))
        return out

    # ---- multi-GPU: one case decomposed over the ranks of a torch.distributed job ----------------
    def ghost_layout(self):
        ng, npat = C.c_int(), C.c_int()
        off, cnt, peer = (np.zeros(64, dtype=np.int32) for _ in range(3))
        self.L.tpp_ghost_layout(self.h, C.byref(ng), C.byref(npat), off.ctypes.data_as(abi.c_int_p), cnt.ctypes.data_as(abi.c_int_p), peer.ctypes.data_as(abi.c_int_p), 64)
        return ng.value, [(int(off[i]), int(cnt[i]), int(peer[i])) for i in range(npat.value)]

    def comm_init_nccl(self):
        """Join the ranks over NCCL (product path): rank 0 creates the unique id, torch.distributed
        broadcasts it, every rank calls tpp_comm_init."""
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        path = nccl_library_path().encode()
        buf = C.create_string_buffer(128)
        if rank == 0 and self.L.tpp_nccl_unique_id(path, buf) != 0:
            self._err("tpp_nccl_unique_id")
        obj = [buf.raw if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        if self.L.tpp_comm_init(self.h, rank, world, obj[0], path) != 0:
            self._err("tpp_comm_init")

    def comm_init_callbacks(self):
        """The same exchange pattern over torch.distributed point-to-point calls on host buffers
        (gloo): used by the CPU tests of the N > 1 path."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        ng, patches = self.ghost_layout()

        def xcb(_user, send, recv, nc):
            s = np.ctypeslib.as_array(send, shape=(ng * nc,))
            r = np.ctypeslib.as_array(recv, shape=(ng * nc,))
            ops, keep = [], []
            for off, cnt, peer in patches:
                ts = torch.from_numpy(s[off * nc : (off + cnt) * nc].copy())
                tr = torch.empty(cnt * nc, dtype=torch.float64)
                keep.append((tr, off, cnt))
                ops += [dist.P2POp(dist.isend, ts, peer), dist.P2POp(dist.irecv, tr, peer)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for tr, off, cnt in keep:
                r[off * nc : (off + cnt) * nc] = tr.numpy()
            return 0

        def rcb(_user, vals, n, op):
            v = np.ctypeslib.as_array(vals, shape=(n,))
            t = torch.from_numpy(v.copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX)
            v[:] = t.numpy()
            return 0

        self._cbs = (XCB(xcb), RCB(rcb))  # keep alive
        if self.L.tpp_comm_callbacks(self.h, rank, world, self._cbs[0], self._cbs[1], None) != 0:
            self._err("tpp_comm_callbacks")

    def load_case_fields(self, case):
        """Start fields from a Case (0/ or the latest time directory: restart)."""
        nC, nF = self.mesh.n_cells, self.mesh.n_faces
        self.set("alpha", case.fields["alpha.water"].internal_array(nC))
        self.set("U", case.fields["U"].internal_array(nC))
        self.set("p_rgh", case.fields["p_rgh"].internal_array(nC))
        self.init_fields()
        # boundary values the next step reads before it re-evaluates them (the old-time wall /
        # atmosphere velocity in ddtCorr, p_rgh on the fixedFluxPressure walls in the non-orthogonal
        # correction): a restart takes them from the time directory, as OpenFOAM does
        for fname, arr, nc in (("U", "U_b", 3), ("p_rgh", "p_rgh_b", 1)):
            cur = self.get(arr).reshape(-1, nc) if nc > 1 else self.get(arr)
            nI, changed = self.mesh.n_internal, False
            for p in self.mesh.patches:
                if p["type"] == "processor" or p["nFaces"] == 0:
                    continue
                v = case.fields[fname].boundary.get(p["name"], {}).get("value")
                if v is None:
                    continue
                v = np.asarray(v, dtype=np.float64)
                sl = slice(p["startFace"] - nI, p["startFace"] - nI + p["nFaces"])
                cur[sl] = v if v.ndim == (1 if nc == 1 else 2) else np.broadcast_to(v, cur[sl].shape)
                changed = True
            if changed:
                self.set(arr, np.ascontiguousarray(cur).reshape(-1))
        if case.restart_delta_t is not None:
            self.set_delta_t(case.restart_delta_t)
