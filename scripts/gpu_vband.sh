#!/usr/bin/env bash
# vband development loop: parity, timing with one phase left out (DCNV3_VBAND_DIAG), optional ncu --set full capture.
set -u
mkdir -p gpurun_out
[ "${PARITY:-1}" = "1" ] && timeout 300 python scripts/vband_check.py 2>&1 | grep -v "^$" | cut -c1-400
for d in ${DIAGS:-1 2 3}; do
  echo "diag=$d"; DCNV3_VBAND_DIAG=$d timeout 120 python scripts/vband_check.py --time-only 2>&1 | grep bfloat16
done
if [ "${NCU:-0}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bwd_vband" -s 4 -c 1 \
      -o gpurun_out/prof_vband -f python scripts/vband_check.py --time-only > gpurun_out/ncu_vband.log 2>&1
  echo "ncu rc=$?"
fi
