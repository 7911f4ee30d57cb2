#!/usr/bin/env bash
# round-2 session-3 loop: quick parity, A/B of the pipelined builders, the training row at the 8-GPU per-rank batch on one GPU
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "${K:-tiled or half or cfg2 or cfg5 or default_shapes}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_quick.log
VAR=DCNV3_SM100_LIB A=$PWD/yolo_somi_b200/libdcnv3_sm100.so B=$PWD/scripts/experiments/diag/libdcnv3_nopipe.so bash scripts/gpu_ab.sh
BENCH_TRAIN_BATCH=16 BENCH_TRAIN_STEPS=10 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_b16.json 2> gpurun_out/bench_b16.err; echo "b16 rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_b16.json').readline()); t = d['train_step']
print('batch 16: %.1f ms/step %.0f img/s; GPU busy %.1f ms' % (t['ms_per_step'], t['img_per_s'], sum(v['ms_per_step'] for v in t['gpu_time_split_rank0'].values())))
PY
