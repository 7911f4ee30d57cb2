#!/usr/bin/env bash
# What bounds the group-slice forward (fwd_gs): diagnostic builds (-DGS_DIAG=n, WRONG results by design) timed by the
# bench.  GS_DIAG=1 keeps the gather's shared-memory loads and drops the arithmetic (the "LSU floor" of DESIGN.md
# section 5, measured instead of argued); GS_DIAG=2 keeps the arithmetic and reads one fixed cell.
# Build first, here (the .so files travel with the snapshot; delete them afterwards): bash scripts/gpu_gsdiag.sh build
set -u
if [ "${1:-}" = build ]; then
  mkdir -p scripts/experiments/diag
  for d in 1 2; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DGS_DIAG=$d -Xcompiler -fPIC,-fvisibility=hidden \
         -Iinclude -Iyolo_somi_b200/csrc -c yolo_somi_b200/csrc/dcnv3_forward_gs.cu -o /tmp/gs_diag$d.o &&
    nvcc -shared -o scripts/experiments/diag/libdcnv3_gsdiag$d.so $(ls yolo_somi_b200/csrc/_obj/*.o | grep -v forward_gs) /tmp/gs_diag$d.o
  done
  exit 0
fi
for d in 0 1 2 0; do
  lib=$PWD/yolo_somi_b200/libdcnv3_sm100.so
  [ $d != 0 ] && lib=$PWD/scripts/experiments/diag/libdcnv3_gsdiag$d.so
  DCNV3_SM100_LIB=$lib python bench.py --steps 30 --warmup 5 --no-cpu --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['passes']
print('GS_DIAG=$d fwd %.1f us bwd %.1f us' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3))"
done
if [ "${NCU:-0}" = 1 ]; then
  for d in 1 2; do
    DCNV3_SM100_LIB=$PWD/scripts/experiments/diag/libdcnv3_gsdiag$d.so ncu --set full --clock-control none -k regex:fwd_gs -s 2 -c 1 \
      -o gpurun_out/prof_gsdiag$d -f python bench.py --steps 2 --warmup 1 --no-cpu --no-train > /dev/null 2>&1
    ncu -i gpurun_out/prof_gsdiag$d.ncu-rep --page raw --csv > gpurun_out/prof_gsdiag${d}_raw.csv
  done
fi
