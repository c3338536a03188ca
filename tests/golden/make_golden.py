#!/usr/bin/env python
"""Regenerates tests/golden/* from the reference checkout (/root/reference, build container
only; the GPU box has no copy).  Everything written here is small and derived from the
reference's committed OpenFOAM-13 run artefacts or from importing
the reference's own Python (`utils/potential_flow.py`, `circularSloshingTank/generate_motion.py`).
"""
import io
import contextlib
import importlib.util
import json
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def read_vtp_arrays(path):
    """The appended-data-free VTP files PyVista wrote for the reference (main.py:773-774): every
    DataArray is base64( UInt32 header [nblocks, blocksize, lastsize, csize...] ) followed by
    base64( zlib blocks ).  stdlib only."""
    import base64
    import re
    import struct
    import zlib

    s = open(path, "rb").read().decode("latin1")
    out = {}
    for m in re.finditer(r"<DataArray ([^>]*)>\s*([^<]*)<", s):
        attrs = dict(re.findall(r'(\w+)="([^"]*)"', m.group(1)))
        if attrs.get("format") != "binary":
            continue
        b64 = m.group(2).strip()
        nb, _, _ = struct.unpack("<3I", base64.b64decode(b64[:16])[:12])
        hlen = (3 + nb) * 4
        hb64 = ((hlen + 2) // 3) * 4
        csizes = struct.unpack(f"<{nb}I", base64.b64decode(b64[:hb64])[12:hlen])
        data = base64.b64decode(b64[hb64:])
        raw, off = b"", 0
        for c in csizes:
            raw += zlib.decompress(data[off : off + c])
            off += c
        dt = {"Float32": "<f4", "Float64": "<f8", "Int64": "<i8", "Int32": "<i4", "UInt8": "u1"}[attrs["type"]]
        a = np.frombuffer(raw, dt)
        nc = int(attrs.get("NumberOfComponents", "1"))
        out[attrs.get("Name", "?")] = a.reshape(-1, nc) if nc > 1 else a
    return out


def g4_series():
    """G4: the reference's 401 committed iso-surfaces (case_..._m0.009/postProcessing/interface/
    interface_t*.vtp, lab frame) reduced frame by frame with the repo's own `iso_wall_mode1` in the
    tank frame (orbit centre from the reference's generate_motion.py law, ramp 2 s: main.py:111-112)."""
    import glob
    import re

    from openfoam_tpp_b200 import interface

    d = os.path.join(REF, "case_H0.208_D0.2_flat_R0.004_f1.88_d20.0_m0.009", "postProcessing", "interface")
    tt = lambda p: float(re.search(r"t([\d.]+)\.vtp", p).group(1))
    rows = []
    for p in sorted(glob.glob(os.path.join(d, "interface_t*.vtp")), key=tt):
        t = tt(p)
        P = read_vtp_arrays(p)["Points"].astype(np.float64)
        tau = min(t / 2.0, 1.0)
        r = 0.004 * (tau * tau * tau * (tau * (tau * 6 - 15) + 10))
        c = (r * np.cos(2 * np.pi * 1.88 * t), r * np.sin(2 * np.pi * 1.88 * t))
        A, ph, z0, n = interface.iso_wall_mode1(P, c, 0.1)
        rows.append((t, A, ph, z0, n, P[:, 2].max(), P[:, 2].min(), P[:, 2].mean(), len(P)))
    return np.array(rows)


def main():
    out = {}
    g4 = g4_series()
    np.savetxt(os.path.join(HERE, "g4_m1_series.csv"), g4, delimiter=",", fmt="%.9g", header="time,A_m1,phase_m1,z0,n_wall_points,max_z,min_z,mean_z,n_points", comments="")
    # G1: post-setFields alpha.water (binary volScalarField written by OpenFOAM 13)
    from openfoam_tpp_b200 import foamfile as ff

    g1 = {}
    for case in ("case_H0.004_D0.0221_flat_R0.005_f2.0", "case_H0.1_D0.02_flat_R0.003_f2.0", "case_H0.208_D0.2_flat_R0.004_f1.88_d20.0_m0.009"):
        f = ff.read_field(os.path.join(REF, case, "0", "alpha.water"))
        a = f.internal
        g1[case] = {"n": int(a.size), "sum": float(a.sum()), "values": sorted(set(np.unique(a).tolist())), "boundary": {k: v.get("type") for k, v in f.boundary.items()}}
    out["G1_alpha"] = g1
    # the smallest one verbatim (62 kB): reader/writer round-trip fixture
    src = os.path.join(REF, "case_H0.004_D0.0221_flat_R0.005_f2.0", "0", "alpha.water")
    with open(src, "rb") as fi, open(os.path.join(HERE, "alpha.water.G1"), "wb") as fo:
        fo.write(fi.read())
    # G2: adaptive time-step sequence (first rows of each probes file) + exact text header
    g2 = {}
    for case in ("case_H0.004_D0.0221_flat_R0.005_f2.0", "case_H0.1_D0.02_flat_R0.003_f2.0", "case_H0.208_D0.2_flat_R0.004_f1.88_d20.0_m0.009", "case_H0.208_D0.2_flat_R0.004_f1.88_d20.0_m0.003"):
        p = os.path.join(REF, case, "postProcessing", "probes", "0", "p")
        with open(p) as f:
            lines = f.readlines()
        times = [float(l.split()[0]) for l in lines[3:]]
        g2[case] = {"head": lines[:8], "n_rows": len(times), "first_times": times[:60], "last_time": times[-1]}
    out["G2_probes"] = g2
    # G3: interface statistics of the 41 895-cell run
    p = os.path.join(REF, "case_H0.208_D0.2_flat_R0.004_f1.88_d20.0_m0.009", "postProcessing", "interface", "interface_summary.csv")
    rows = np.loadtxt(p, delimiter=",", skiprows=1)
    out["G3_interface"] = {"n": int(rows.shape[0]), "t": rows[:, 0].tolist(), "max_z": rows[:, 1].tolist(), "min_z": rows[:, 2].tolist(), "mean_z": rows[:, 3].tolist()}
    # G5: analytic potential flow (imported reference code)
    pf = load(os.path.join(REF, "utils", "potential_flow.py"), "potential_flow")
    with contextlib.redirect_stdout(io.StringIO()):
        omegas, eps = pf.compute_natural_frequencies(0.1, 0.104)[:2] if isinstance(pf.compute_natural_frequencies(0.1, 0.104), tuple) else (pf.compute_natural_frequencies(0.1, 0.104), None)
    out["G5_potential"] = {"R": 0.1, "d": 0.104, "omega_1n": np.asarray(omegas).tolist()}
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            A = pf.compute_wall_amplitude(0.1, 0.004, 2 * np.pi * 1.88, 0.104)
        out["G5_potential"]["A_PT"] = float(A if np.isscalar(A) else A[0])
    except Exception as e:  # signature differs: record what we can
        out["G5_potential"]["A_PT_error"] = str(e)
    # motion table: the reference generator's own output for the cfg1 parameters (first rows)
    gm = load(os.path.join(REF, "circularSloshingTank", "generate_motion.py"), "generate_motion")
    tmp = os.path.join(HERE, "_tmp_6DoF.dat")
    gm.generate_motion(0.005, 2.0, 0.05, 0.001, 0.02, tmp)
    with open(tmp) as f:
        out["motion_table_text"] = f.read()
    os.remove(tmp)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote golden.json", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
