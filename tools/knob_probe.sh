#!/bin/sh
# step time and the main V-cycle kernels under knob settings given as arguments (tuning aid):
#   sh tools/knob_probe.sh "X=0" "TPP_ELLC_OV=6" ...
for cfg in "$@"; do
  env $cfg python bench.py --steps 10 --warmup 3 --spinup 10 --no-cpu --kernel-table /tmp/kt.json > /tmp/b.json 2>/dev/null
  python - "$cfg" <<'PY'
import json, sys
k = json.load(open('/tmp/kt.json')); b = json.load(open('/tmp/b.json'))
t = {r[0]: r for r in k['kernels_launches_ms_GBps']}
def us(n): return f"{n} {1e3 * t[n][2] / t[n][1]:.1f}us x{t[n][1] // 2}" if n in t else ""
print(f"{sys.argv[1]:34s} step {b['ms_per_step']:.2f} ms  iters {b['config']['iters_mean']}  " + "  ".join(us(n) for n in ("v_tail", "v_jacobi_csr", "v_jacobi_first_csr", "v_jacobi_corr_csr", "v_residual_csr", "U_recon", "HbyA", "grad_U", "mom_face", "alpha_flux", "grad_scalar")))
PY
done
