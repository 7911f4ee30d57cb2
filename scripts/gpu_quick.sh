#!/usr/bin/env bash
# Quick GPU loop: selected parity tests (K=pytest -k expr) + a short bench.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "${K:-tiled or half or cfg2}" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -${TAIL:-15} gpurun_out/pytest_quick.log
python bench.py --steps 20 --warmup 5 --no-cpu --no-train > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_quick.json').readline()); p = d['passes']
    print('fwd %.1f us  bwd %.1f us  step %.1f us  pts/s %.3g  frac %.3f  e2e %.2f ms' % (p['fwd_ms']*1e3, p['bwd_ms']*1e3, d['ms_per_step']*1e3, d['value'], p['step_frac'], d['e2e']['ms_per_step']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
