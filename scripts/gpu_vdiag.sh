#!/usr/bin/env bash
# What each phase of bwd_vmma costs: diagnostic builds (-DVMMA_DIAG=n, WRONG results by design) timed by the bench.
# Build them first (here, not on the GPU box; the .so files travel with the snapshot and are deleted afterwards):
#   cd yolo_somi_b200/csrc && for d in 1 2 3 4; do
#     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DVMMA_DIAG=$d -Xcompiler -fPIC,-fvisibility=hidden \
#          -I../../include -I. -c dcnv3_backward_vmma.cu -o /tmp/vmma_diag$d.o &&
#     nvcc -shared -o ../../scripts/experiments/diag/libdcnv3_diag$d.so $(ls _obj/*.o | grep -v backward_vmma) /tmp/vmma_diag$d.o; done
set -u
for d in 0 1 2 3 4 0; do
  lib=$PWD/yolo_somi_b200/libdcnv3_sm100.so
  [ $d != 0 ] && lib=$PWD/scripts/experiments/diag/libdcnv3_diag$d.so
  DCNV3_SM100_LIB=$lib python bench.py --steps 30 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['passes']
print('DIAG=$d fwd %.1f us bwd %.1f us step %.1f us' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3, d['ms_per_step']*1e3))"
done
