#!/bin/bash
# scratch A/B template for one GPU call (edit the environment switches per experiment)
mkdir -p gpurun_out
for cfg in "1 1" "0 0" "1 1" "0 0"; do
set -- $cfg
GNAS_SIDE_PRIO=$1 GNAS_CAPTURE_PRIO=$2 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_ab.log 2> gpurun_out/bench.err; echo "prio=$1 cap=$2 rc=$?"; tail -1 gpurun_out/bench_ab.log | cut -c60-190; tail -2 gpurun_out/bench.err
done
GNAS_SIDE_PRIO=1 GNAS_CAPTURE_PRIO=1 python tools/timeline_step.py 2>&1 | tail -6
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
