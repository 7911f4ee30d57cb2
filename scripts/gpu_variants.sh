#!/usr/bin/env bash
# bench the kernel variants selected by the development env knobs; outputs -> gpurun_out/variants.txt
mkdir -p gpurun_out; : > gpurun_out/variants.txt
run() { echo "## $*" >> gpurun_out/variants.txt; env "$@" python bench.py --steps 20 --warmup 5 --no-cpu 2>>gpurun_out/variants.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['passes']
print('fwd %.1f us  bwd %.1f us  step %.1f us  pts/s %.3g  frac %.3f' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3, d['ms_per_step']*1e3, d['value'], p['step_frac']))" >> gpurun_out/variants.txt; }
for v in ${VARIANTS:-"DCNV3_NV=1 DCNV3_WEIGHTS=split DCNV3_BWD_PAIR=1"}; do run $(echo $v | tr "," " "); done
cat gpurun_out/variants.txt
