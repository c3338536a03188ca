#!/usr/bin/env python
"""Step-by-step diagnostics around an instability of the cfg4 validation run."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.argv_backup = sys.argv; 
import numpy as np
import bench
from openfoam_tpp_b200 import meshgen, motion, case as cs, foamfile as ff, solver as sv
import tempfile
ap = argparse.ArgumentParser(); ap.add_argument('--t0', type=float, default=5.9); ap.add_argument('--t1', type=float, default=7.0)
ap.add_argument('--rings', type=int, default=10); ap.add_argument('--layers', type=int, default=22); ap.add_argument('--out', default='gpurun_out/blowup')
a = ap.parse_args()
C = bench.CASE
mesh = meshgen.cylinder_mesh(C['H'], C['D'], a.rings, a.layers, 'flat', 'tet')
with tempfile.TemporaryDirectory() as tmp:
    cs.write_template(tmp, end_time=a.t1, write_interval=0.05, fill_z=C['H']/2)
    motion.write_table(os.path.join(tmp,'constant','6DoF.dat'), motion.orbital_table(C['R'], C['freq'], 21.0, C['dt'], C['ramp']))
    cfg = cs.read_config(tmp, None)
    fields = {n: ff.read_field(os.path.join(tmp,'0',n)) for n in ('U','alpha.water','p_rgh')}
    cs._bc_tables(cfg, mesh, fields, '0')
cfg.start_time = 0.0
s = sv.Solver(mesh, cfg); s.set('alpha', bench.initial_alpha(mesh)); s.init_fields()
V = s.get('V'); vol0 = (s.get('alpha')*V).sum()
while s.info()['t'] < a.t0 - 1e-9:
    if s.run_to_write(10**9) != 1: break
print('reached', s.info()['t'], 'step', int(s.info()['step']), 'vol', (s.get('alpha')*V).sum()/vol0, flush=True)
names = ['alpha','U','p_rgh','phi','Uf','alpha_b','U_b','p_rgh_b','pGrad_b','rho','rho_b','p']
prev = None; dumped = False; log = open(a.out + '_log.csv', 'w')
log.write('step,t,dt,Co,alphaCo,it0,r0,it1,r1,min_rAU,n_neg,maxU,vol,amax,amin\n')
while s.info()['t'] < a.t1:
    state = {n: s.get(n) for n in names}; inf0 = s.info()
    s.step(1)
    i = s.info(); rAU = s.get('rAU'); U = s.get('U'); al = s.get('alpha')
    vol = (al*V).sum()/vol0; mu = np.abs(U).max()
    log.write(f"{int(i['step'])},{i['t']:.9g},{i['dt']:.4e},{i['Co']:.4f},{i['alphaCo']:.4f},{int(i['it0'])},{i['r0']:.2e},{int(i['it1'])},{i['r1']:.2e},{rAU.min():.3e},{int((rAU<0).sum())},{mu:.4g},{vol:.8f},{al.max():.6f},{al.min():.3e}\n"); log.flush()
    if not dumped and (mu > 30 or vol < 0.99 or not np.isfinite(mu)):
        np.savez_compressed(a.out + '_state.npz', t=inf0['t'], dt=inf0['dt'], step=inf0['step'], **state)
        dumped = True; print('dumped state before step', int(i['step']), 't', inf0['t'], 'maxU', mu, 'vol', vol, flush=True)
    if dumped and i['step'] > inf0['step'] + 300: break
print('end', s.info()['t'])
