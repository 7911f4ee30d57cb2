#!/usr/bin/env bash
# A/B of one environment knob on the bench: VAR=name A=value B=value bash scripts/gpu_ab.sh
set -u
for z in "$A" "$B" "$A" "$B"; do
  env "$VAR=$z" python bench.py --steps 30 --warmup 5 --no-cpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['passes']
print('$VAR=$z fwd %.1f us bwd %.1f us step %.1f us' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3, d['ms_per_step']*1e3))"
done
