#!/bin/sh
# AddressSanitizer pass over the host emulation of the kernels (TPP_EMU build of the same source):
# catches out-of-bounds indexing in the kernel bodies and in the host driver without a GPU.
#   sh tools/asan_check.sh        (from the repo root; ~2 min)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${TMPDIR:-/tmp}/tpp_asan
mkdir -p "$OUT"
g++ -x c++ -DTPP_EMU -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -fPIC -ffp-contract=off -shared \
    -o "$OUT/libtppvof_asan.so" "$ROOT/openfoam-tpp_b200/csrc/tppvof.cu"
export ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
cd "$ROOT"
python - "$OUT" <<'PY'
import sys, textwrap
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from openfoam_tpp_b200 import solver as sv, case as cs
out = sys.argv[1]; LIB = out + '/libtppvof_asan.so'
for cell, geo in (("tet", "flat"), ("prism", "cap")):
    d = f'{out}/case_{cell}'
    cs.setup_case(d, H=0.004, D=0.0221, R=0.005, freq=2.0, duration=1.0, n_rings=6, n_layers=6, cell=cell, geo=geo)
    c = cs.Case(d); g = sv.Solver(c.mesh, c.cfg, lib_path=LIB); g.load_case_fields(c)
    g.set_probes([g.find_cell(np.array([0.001, 0.001, 0.001])), -1]); g.step(3); g.close()
d = f'{out}/tutorial'; cs.setup_tutorial_case(d, nx=6, ny=12, nz=9, end_time=1.0)
c = cs.Case(d); g = sv.Solver(c.mesh, c.cfg, lib_path=LIB); g.load_case_fields(c); g.step(3); g.close()
import test_decomposed as t
open(out + '/worker.py', 'w').write(textwrap.dedent(t.WORKER.format(root='.', lib=LIB, nr=6, nl=12, steps=2, sigma=0.0)))
print('single-rank cases clean')
PY
# the case reader / writer behind tpp_open (csrc/tpp_caseio.h): whole runs, resume, and every input file truncated
python - "$OUT" <<'PY'
import os, shutil, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from openfoam_tpp_b200 import solver as sv, case as cs, foamfile as ff
import test_caseio as t
out = sys.argv[1]; LIB = out + '/libtppvof_asan.so'
for binary in (True, False):
    d = f'{out}/cio_{int(binary)}'; shutil.rmtree(d, ignore_errors=True); t._setup(d)
    t._set_entry(d + '/system/controlDict', 'writeFormat', 'binary' if binary else 'ascii')
    t._set_entry(d + '/system/controlDict', 'endTime', '0.006')
    if not binary:
        ff.write_polymesh(d, ff.read_polymesh(d), binary=False)
    s = sv.Solver.open(d, lib_path=LIB); s.run_case(interface=True); s.close()
    s = sv.Solver.open(d, lib_path=LIB); assert s.case_query('start_time') == '0.006'; s.write_time(); s.close()
    for f in ('constant/polyMesh/owner', 'constant/polyMesh/faces', 'constant/polyMesh/points', 'constant/polyMesh/boundary', 'system/fvSolution', 'constant/6DoF.dat', '0.006/alpha.water', '0.006/phi', '0.006/uniform/time'):
        p = os.path.join(d, f); raw = open(p, 'rb').read()
        for cut in (len(raw) // 2, len(raw) // 3, 40, 0):
            open(p, 'wb').write(raw[:cut])
            try:
                sv.Solver.open(d, lib_path=LIB).close()
            except sv.SolverError:
                pass
        open(p, 'wb').write(raw)
print('case directories clean')
PY
TPP_TAIL_ROWS=300 TPP_COARSEST=100 python -m torch.distributed.run --nnodes=1 --nproc-per-node=3 --master-addr 127.0.0.1 --master-port 29643 "$OUT/worker.py" 2>&1 | grep "RANK. OK"
echo "asan: clean"
