#!/bin/bash
# one GPU call: tests, smoke, bench, ncu launch list (run from the repo root on the GPU box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400; tail -3 gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
  timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-graph > gpurun_out/plain_ng.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
  echo "ncu rc=$?"
fi
