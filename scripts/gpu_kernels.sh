#!/usr/bin/env bash
# durations of the full-batch launches of the sampling kernels under the given env: bash scripts/gpu_kernels.sh VAR=val ...
set -u; mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 1 --no-cpu"
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fwd_|bwd_|narrow" -c 24 --csv --log-file gpurun_out/launches_k.csv $CMD > gpurun_out/ncu_k.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.DictReader(l for l in open('gpurun_out/launches_k.csv') if not l.startswith('=='))]
agg = collections.defaultdict(list)
for r in rows:
    agg[r['Kernel Name'][:44]].append(float(r['Metric Value'].replace(',', '')) / 1e3)
for k, v in agg.items():
    print(f"{k:46s} n={len(v):2d}  median {sorted(v)[len(v)//2]:7.1f} us   min {min(v):7.1f}")
PY
