"""TEST INFRASTRUCTURE: ctypes wrapper of oracle/libvoforacle.so (the CPU restatement).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from openfoam_tpp_b200 import abi  # noqa: E402  (struct layouts only)

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libvoforacle.so")
    src = os.path.join(_HERE, "vof_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libvoforacle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.MeshStruct), C.POINTER(abi.ConfigStruct)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        for f in (L.orc_get, L.orc_set):
            f.restype = C.c_long
            f.argtypes = [C.c_void_p, C.c_char_p, abi.c_double_p, C.c_long]
        L.orc_size.restype = C.c_long
        L.orc_size.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_get_int.restype = C.c_long
        L.orc_get_int.argtypes = [C.c_void_p, C.c_char_p, abi.c_int_p, C.c_long]
        L.orc_stage.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_run_to_write.argtypes = [C.c_void_p, C.c_long]
        L.orc_info.argtypes = [C.c_void_p, abi.c_double_p]
        L.orc_solve.argtypes = [C.c_void_p, C.POINTER(abi.SolverStruct)] + [abi.c_double_p] * 6
        L.orc_set_probes.argtypes = [C.c_void_p, C.c_int, abi.c_int_p]
        L.orc_probe_log.restype = C.c_long
        L.orc_probe_log.argtypes = [C.c_void_p, abi.c_double_p, C.c_long]
        L.orc_find_cell.argtypes = [C.c_void_p, abi.c_double_p]
        _LIB = L
    return _LIB


class Oracle:
    def __init__(self, mesh, cfg):
        self.L = lib()
        m, c, self._keep = abi.build_structs(mesh, cfg)
        self.h = self.L.orc_create(C.byref(m), C.byref(c))
        if not self.h:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())
        self.mesh, self.cfg = mesh, cfg

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def get(self, name):
        n = self.L.orc_size(self.h, name.encode())
        if n < 0:
            raise KeyError(name)
        a = np.empty(n, dtype=np.float64)
        self.L.orc_get(self.h, name.encode(), a.ctypes.data_as(abi.c_double_p), n)
        return a

    def set(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        if self.L.orc_set(self.h, name.encode(), a.ctypes.data_as(abi.c_double_p), a.size) < 0:
            raise KeyError(name)

    def get_int(self, name):
        n = self.L.orc_get_int(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        a = np.empty(n, dtype=np.int32)
        self.L.orc_get_int(self.h, name.encode(), a.ctypes.data_as(abi.c_int_p), n)
        return a

    def stage(self, name):
        if self.L.orc_stage(self.h, name.encode()) != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())

    def step(self, n=1):
        self.L.orc_step(self.h, n)

    def run_to_write(self, max_steps=10**9):
        return self.L.orc_run_to_write(self.h, max_steps)

    def info(self):
        o = np.zeros(16)
        self.L.orc_info(self.h, o.ctypes.data_as(abi.c_double_p))
        keys = ["t", "dt", "step", "Co", "alphaCo", "it0", "r00", "r0", "it1", "r01", "r1", "refCell", "deltaN", "writeIndex", "levels0", "levels1"]
        return dict(zip(keys, o))

    def solve(self, ctl, diag, upper, b, x0=None):
        x = np.zeros(self.mesh.n_cells) if x0 is None else np.array(x0, dtype=np.float64)
        r0, r = C.c_double(), C.c_double()
        s = abi.solver_struct(ctl)
        dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(abi.c_double_p)
        d_, u_, b_ = (np.ascontiguousarray(v, dtype=np.float64) for v in (diag, upper, b))
        it = self.L.orc_solve(self.h, C.byref(s), d_.ctypes.data_as(abi.c_double_p), u_.ctypes.data_as(abi.c_double_p), b_.ctypes.data_as(abi.c_double_p), x.ctypes.data_as(abi.c_double_p), C.cast(C.byref(r0), abi.c_double_p), C.cast(C.byref(r), abi.c_double_p))
        return x, it, r0.value, r.value

    def set_probes(self, cells):
        a = np.ascontiguousarray(cells, dtype=np.int32)
        self._nprobe = a.size
        self.L.orc_set_probes(self.h, a.size, a.ctypes.data_as(abi.c_int_p))

    def probe_log(self, cap=1 << 20):
        w = 1 + self._nprobe
        a = np.empty((cap, w))
        n = self.L.orc_probe_log(self.h, a.ctypes.data_as(abi.c_double_p), cap)
        return a[:n].copy()

    def find_cell(self, xyz):
        a = np.ascontiguousarray(xyz, dtype=np.float64)
        return self.L.orc_find_cell(self.h, a.ctypes.data_as(abi.c_double_p))

    def load_case_fields(self, case):
        """Initial fields from a Case (0/ or the latest time directory)."""
        nC = self.mesh.n_cells
        self.set("alpha", case.fields["alpha.water"].internal_array(nC))
        self.set("U", case.fields["U"].internal_array(nC))
        self.set("p_rgh", case.fields["p_rgh"].internal_array(nC))
        self.stage("alphaBCs")
        self.stage("mixture")
