mkdir -p gpurun_out
for k in 1 0 1 0; do
GNAS_WG_LATE=$k timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_ab.log 2> gpurun_out/bench.err; echo "late=$k rc=$?"; tail -1 gpurun_out/bench_ab.log | cut -c60-190
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
