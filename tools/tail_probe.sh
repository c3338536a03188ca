#!/bin/sh
# v_tail time per launch under a few knob settings (tuning aid): sh tools/tail_probe.sh
for cfg in "X=0" "TPP_CITER=1" "TPP_TAIL_CTAS=32" "TPP_TAIL_CTAS=74" "TPP_TAIL_NPRE=1 TPP_TAIL_NPOST=1" "TPP_TAIL_ELL=0" "TPP_CG_SMEM=0"; do
  env $cfg python bench.py --steps 5 --warmup 3 --spinup 5 --no-cpu --kernel-table /tmp/kt.json > /tmp/b.json 2>/dev/null
  python - "$cfg" <<'PY'
import json, sys
k = json.load(open('/tmp/kt.json')); b = json.load(open('/tmp/b.json'))
t = [r for r in k['kernels_launches_ms_GBps'] if r[0] == 'v_tail'][0]
print(f"{sys.argv[1]:40s} v_tail {1e3 * t[2] / t[1]:7.1f} us/launch x {t[1] // 2}/step   step {b['ms_per_step']:.2f} ms  iters {b['config']['iters_mean']}")
PY
done
