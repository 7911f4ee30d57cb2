#!/usr/bin/env bash
# Programmatic dependent launch between bwd_dots -> bwd_vmma -> narrow (default) against full serialisation (DCNV3_PDL=0).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tiled and default or cta_forms or cfg2 or fp16 or inside_their or capture or determin or graph" 2>&1 | tail -3
for z in 0 1 0 1; do
  DCNV3_PDL=$z python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['passes']
print('PDL=$z fwd %.1f us bwd %.1f us step %.1f us e2e %.2f ms' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3, d['ms_per_step']*1e3, d['e2e']['ms_per_step']))"
done
