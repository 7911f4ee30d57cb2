#!/usr/bin/env bash
# whole-step CUDA graph: the GPU test, then the training row at the 8-GPU per-rank batch (16) on one GPU, eager vs graph
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step.py -m gpu -x -q > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_graph.log
BENCH_TRAIN_BF16=0 BENCH_TRAIN_BATCH=${B:-16} BENCH_TRAIN_STEPS=10 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_graph.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_graph.json').readline()); t = d['train_step']
print('headline %.1f img/s (%s)' % (t['img_per_s'], t.get('mode')))
print('eager', t.get('eager'))
print('graph', json.dumps(t.get('whole_step_cuda_graph'))[:1500])
PY
