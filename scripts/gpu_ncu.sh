#!/usr/bin/env bash
# env "$@" ncu --set full of selected kernels: KREGEX='fwd_tile' OUT=prof_x bash scripts/gpu_ncu.sh [env...]
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-train"
env "$@" $CMD > gpurun_out/plain.log 2>&1 &&
env "$@" ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-fwd_|bwd_}" -s ${SKIP:-4} -c ${COUNT:-2} \
    -o gpurun_out/${OUT:-prof} -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
