#!/usr/bin/env bash
# Four groups per 256-thread CTA (default for gc == 16) against eight per 512-thread CTA (DCNV3_GS_KG=8).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "cta_forms or gc32 or inside_their" 2>&1 | tail -3
for kg in 8 4 8 4; do
  DCNV3_GS_KG=$kg python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['passes']
print('KG=$kg fwd %.1f us bwd %.1f us step %.1f us' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3, d['ms_per_step']*1e3))"
done
