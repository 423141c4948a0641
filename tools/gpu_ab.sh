mkdir -p gpurun_out
for cfg in "1 1" "0 1" "1 0" "0 0" "1 1"; do
set -- $cfg
GNAS_PRE_SIDE=$1 GNAS_C15DX_SIDE=$2 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_ab.log 2> gpurun_out/bench.err; echo "pre=$1 c15dx=$2 rc=$?"; tail -1 gpurun_out/bench_ab.log | cut -c60-190
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
