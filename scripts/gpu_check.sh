#!/usr/bin/env bash
# One gpurun call: GPU parity tests, smoke, a short bench, then (only if the plain bench exited 0)
# the ncu launch list and one full capture of the two sampling kernels.  Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -25 gpurun_out/pytest.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  python bench.py --steps 2 --warmup 1 --no-cpu --no-train > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-train > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:"fwd_|bwd_|far_points" -s 6 -c 6 \
      -o gpurun_out/prof_r2 -f python bench.py --steps 2 --warmup 1 --no-cpu --no-train > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
