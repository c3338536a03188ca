#!/usr/bin/env python
"""Integration run: the D = 0.2 m tank (case_H0.208_D0.2_flat_R0.004_f1.88, 20 s in the
reference) stepped by the GPU solver or by the CPU oracle, recording every `--every` seconds the
interface statistics the reference's post-processing reports (max/min/mean elevation) and the
first azimuthal mode at the wall (amplitude, phase) in the tank frame.

  python tools/validate_run.py --impl gpu|oracle --rings 10 --layers 22 --end 20 --out series.csv
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402  (workload definition shared with bench.py)
from openfoam_tpp_b200 import interface, meshgen, motion  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="gpu")
    ap.add_argument("--rings", type=int, default=10)
    ap.add_argument("--layers", type=int, default=22)
    ap.add_argument("--end", type=float, default=20.0)
    ap.add_argument("--every", type=float, default=0.05)
    ap.add_argument("--out", default="series.csv")
    ap.add_argument("--tight", action="store_true", help="solve p_rgh to 1e-12 instead of the reference's tolerances")
    ap.add_argument("--mesh", default="structured", choices=("structured", "unstructured"), help="unstructured: Delaunay tets of size --lc (the gmsh stand-in), cells in triangulator order")
    ap.add_argument("--lc", type=float, default=0.009, help="generate_mesh.py's MeshSize (main.py:113; the golden case m0.009)")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    C = bench.CASE
    if a.mesh == "unstructured":
        mesh = meshgen.unstructured_cylinder_mesh(C["H"], C["D"], a.lc, seed=a.seed, iters=40)
    else:
        mesh = meshgen.cylinder_mesh(C["H"], C["D"], a.rings, a.layers, "flat", "tet")
    import tempfile

    from openfoam_tpp_b200 import case as cs
    from openfoam_tpp_b200 import foamfile as ff

    with tempfile.TemporaryDirectory() as tmp:
        cs.write_template(tmp, end_time=a.end, write_interval=a.every, fill_z=C["H"] / 2)
        rows = motion.orbital_table(C["R"], C["freq"], a.end, C["dt"], C["ramp"])
        motion.write_table(os.path.join(tmp, "constant", "6DoF.dat"), rows)
        cfg = cs.read_config(tmp, None)
        fields = {n: ff.read_field(os.path.join(tmp, "0", n)) for n in ("U", "alpha.water", "p_rgh")}
        cs._bc_tables(cfg, mesh, fields, "0")
    cfg.start_time = 0.0
    if a.tight:
        for s in (cfg.p_rgh, cfg.p_rgh_final):
            s.tolerance, s.rel_tol, s.max_iter = 1e-12, 0.0, 300
    if a.impl == "gpu":
        from openfoam_tpp_b200 import solver as sv

        s = sv.Solver(mesh, cfg)
        s.set("alpha", bench.initial_alpha(mesh))
        s.init_fields()
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle

        s = oracle.Oracle(mesh, cfg)
        s.set("alpha", bench.initial_alpha(mesh))
        s.stage("alphaBCs")
        s.stage("mixture")
    cols = interface.ColumnSampler(mesh) if a.mesh == "structured" else None
    V = meshgen.cell_geometry(mesh)[1]
    edges = interface.mesh_edges(mesh)  # the reference's own metric: points of the alpha = 0.5 iso-surface (main.py:751-780)
    t0 = time.perf_counter()
    with open(a.out, "w") as f:
        f.write("time,max_z,min_z,mean_z,A_m1,phase_m1,alpha_volume,step,it_final,res_final,wall_s,iso_max_z,iso_min_z,iso_mean_z,iso_points,iso_A_m1,iso_phase_m1\n")

        def rec():
            al = s.get("alpha")
            i = s.info()
            # tank frame: z statistics do not see the horizontal translation, and the m = 1 fit
            # wants the tank-frame azimuth (the golden G4 series subtracts the orbit centre)
            iso = interface.iso_points(mesh, mesh.points, interface.cell_to_point(mesh, al), 0.5, edges)
            z = iso[:, 2] if len(iso) else np.zeros(1)
            iA, iph, _, _ = interface.iso_wall_mode1(iso, (0.0, 0.0), 0.5 * C["D"])
            if cols is not None:
                mx, mn, me = cols.summary(al)
                A, ph, _ = cols.wall_mode1(al)
            else:
                mx, mn, me, A, ph = z.max(), z.min(), z.mean(), iA, iph
            f.write(f"{i['t']:.9g},{mx:.9g},{mn:.9g},{me:.9g},{A:.9g},{ph:.9g},{float((al * V).sum()):.15g},{int(i['step'])},{int(i['it1'])},{i['r1']:.3e},{time.perf_counter() - t0:.2f},{z.max():.9g},{z.min():.9g},{z.mean():.9g},{len(iso)},{iA:.9g},{iph:.9g}\n")
            f.flush()

        rec()
        while s.run_to_write(10**9) == 1:
            rec()
    i = s.info()
    print(f"{a.impl}: {mesh.n_cells} cells, {int(i['step'])} steps to t={i['t']:.4g} in {time.perf_counter() - t0:.1f} s")


if __name__ == "__main__":
    main()
