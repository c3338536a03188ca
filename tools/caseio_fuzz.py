"""Mutation fuzzing of every input file of a case directory against tpp_open (csrc/tpp_caseio.h) on an
AddressSanitizer + UBSan build of the host emulation: a damaged file must be an error code (or a case that
still opens), never a stray read.  Run after `sh tools/asan_check.sh` (which builds the library):
    LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
        SEED=7 [HOST=python] python tools/caseio_fuzz.py
Found so far: face offsets overrunning the point labels (fixed in the reader and in abi.build_structs)."""
import os, random, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from openfoam_tpp_b200 import solver as sv, foamfile as ff
import test_caseio as t
OUT = os.path.join(os.environ.get('TMPDIR', '/tmp'), 'tpp_asan')
LIB = OUT + '/libtppvof_asan.so'
random.seed(int(os.environ.get("SEED", "7")))
n_ok = n_err = 0
for binary in (False, True):
    d = f'{OUT}/fuzz_{int(binary)}'; shutil.rmtree(d, ignore_errors=True); t._setup(d)
    t._set_entry(d + '/system/controlDict', 'writeFormat', 'binary' if binary else 'ascii')
    t._set_entry(d + '/system/controlDict', 'endTime', '0.003')
    if not binary:
        ff.write_polymesh(d, ff.read_polymesh(d), binary=False)
    s = sv.Solver.open(d, lib_path=LIB); s.run_case(); s.close()
    files = ['constant/polyMesh/owner', 'constant/polyMesh/neighbour', 'constant/polyMesh/faces', 'constant/polyMesh/points', 'constant/polyMesh/boundary', 'constant/polyMesh/cellZones',
             'system/fvSolution', 'system/fvSchemes', 'system/controlDict', 'system/functions', 'constant/dynamicMeshDict', 'constant/6DoF.dat', 'constant/g', 'constant/phaseProperties',
             '0.003/alpha.water', '0.003/U', '0.003/p_rgh', '0.003/phi', '0.003/Uf', '0.003/uniform/time']
    files = [f for f in files if os.path.exists(os.path.join(d, f))]
    for it in range(260):
        f = random.choice(files); p = os.path.join(d, f); raw = open(p, 'rb').read()
        b = bytearray(raw)
        hdr_end = raw.find(b'}') + 1
        kind = random.randrange(5)
        pos = random.randrange(hdr_end, max(hdr_end + 1, len(b)))
        if kind == 0:
            b[pos] = random.choice(b'(){};"0123456789 \n/*#$aZ.-e')
        elif kind == 1:
            del b[pos:pos + random.randrange(1, 40)]
        elif kind == 2:
            b[pos:pos] = bytes(random.choice(b'(){};" \n0123456789.e-') for _ in range(random.randrange(1, 12)))
        elif kind == 3:
            k = random.randrange(hdr_end, max(hdr_end + 1, len(b))); b[pos:pos] = b[k:k + random.randrange(1, 200)]
        else:  # corrupt a count
            import re
            m = list(re.finditer(rb'\n(\d+)\n\(', raw))
            if m:
                mm = random.choice(m); b = bytearray(raw[:mm.start(1)] + str(random.choice([0, 1, int(mm.group(1)) + 1, int(mm.group(1)) - 1, 10**12, 2**31])).encode() + raw[mm.end(1):])
        open(p, 'wb').write(bytes(b))
        try:
            if os.environ.get('HOST', 'library') == 'python':  # the Python reader in front of tpp_create
                from openfoam_tpp_b200 import case as cs
                c = cs.Case(d); h = sv.Solver(c.mesh, c.cfg, lib_path=LIB); h.load_case_fields(c)
            else:
                h = sv.Solver.open(d, lib_path=LIB)
            h.close(); n_ok += 1
        except Exception as e:  # noqa: BLE001 - any refusal is fine; only a sanitizer abort is a finding
            n_err += 1
        open(p, 'wb').write(raw)
print('fuzz done: accepted', n_ok, 'refused', n_err)
