mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c60-190
python tools/timeline_step.py > gpurun_out/r02_timeline.txt 2>&1; tail -11 gpurun_out/r02_timeline.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
