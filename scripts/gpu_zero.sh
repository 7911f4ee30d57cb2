#!/usr/bin/env bash
# Plane zeroed inside bwd_dots (default) against the memset on a side stream (DCNV3_ZERO=side).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tiled and default or cta_forms or cfg2 or fp16 or inside_their or capture or determin" 2>&1 | tail -3
for z in side dots side dots; do
  DCNV3_ZERO=$z python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['passes']
print('ZERO=$z fwd %.1f us bwd %.1f us step %.1f us' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3, d['ms_per_step']*1e3))"
done
