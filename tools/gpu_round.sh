#!/bin/bash
# one GPU call: tests, smoke, bench, (ncu) launch list, gradient-error distribution (run from the repo root on the GPU box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400; tail -3 gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
  (for tag in full pdarts3; do TOP=6 timeout 300 python tools/full_grad_err.py $tag; done) > gpurun_out/r02_full_grad_err.txt 2>&1; tail -8 gpurun_out/r02_full_grad_err.txt
  timeout 900 python bench.py --workload pdarts3 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_pdarts3.log 2> gpurun_out/bench.err; echo "pdarts3 rc=$?"; tail -1 gpurun_out/bench_pdarts3.log | cut -c1-300
  timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline --no-graph > gpurun_out/plain_ng.log 2>&1 && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager-baseline --no-graph > gpurun_out/ncu_launch.log 2>&1
  echo "ncu rc=$?"
fi
