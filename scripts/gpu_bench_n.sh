#!/usr/bin/env bash
# N-rank bench with the graph leg: GPUS=2 bash scripts/gpu_bench_n.sh
set -u
mkdir -p gpurun_out
N=${GPUS:-2}
BENCH_TRAIN_GRAPH_TIMEOUT=240 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$N.json').readline()); t = d['train_step']
print('value %.3g pts/s  step %.1f us' % (d['value'], d['ms_per_step']*1e3))
print('headline %.1f img/s (%s)' % (t['img_per_s'], t.get('mode')))
print('eager', t.get('eager'))
g = t.get('whole_step_cuda_graph') or {}
print('graph', {k: g.get(k) for k in ('img_per_s', 'ms_per_step', 'cuda_graph', 'error', 'limiter')})
print('nccl', (g.get('gpu_time_split_rank0') or {}).get('nccl'), 'eager nccl', t['gpu_time_split_rank0'].get('nccl'))
print('weak', t.get('weak_scaling_128_per_gpu'))
PY
