#!/usr/bin/env bash
# A/B of a variant library against the default one on the op bench, with the value-kernel parity tests on the variant:
#   LIB=scripts/experiments/diag/libdcnv3_x.so bash scripts/gpu_ab_lib.sh
set -u
DCNV3_SM100_LIB=$PWD/$LIB timeout 600 python -m pytest tests/test_dcnv3_gpu.py -m gpu -x -q -k "${K:-cfg2 or default_kernels_fp16 or half_precision}" 2>&1 | tail -2
VAR=DCNV3_SM100_LIB A=$PWD/yolo_somi_b200/libdcnv3_sm100.so B=$PWD/$LIB bash scripts/gpu_ab.sh
