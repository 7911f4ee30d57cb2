#!/usr/bin/env bash
# per-kernel durations (ncu launch list) of a short bench run under the given env: bash scripts/gpu_launches.sh VAR=val ...
set -u; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu"
env "$@" $CMD > gpurun_out/plain.log 2>&1 &&
env "$@" ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_x.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.DictReader(l for l in open('gpurun_out/launches_x.csv') if not l.startswith('=='))]
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows:
    agg[r['Kernel Name'][:48]][r['Metric Name']].append(float(r['Metric Value'].replace(',', '')))
for k, m in agg.items():
    print(k)
    for name, v in m.items():
        print(f"    {name:70s} {sum(v)/len(v):16.1f}  (n={len(v)})")
PY
