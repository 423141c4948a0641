mkdir -p gpurun_out
for k in 1 0 1 0; do
GNAS_WG_AUTO=$k timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_ab.log 2> gpurun_out/bench.err; echo "auto=$k rc=$?"; tail -1 gpurun_out/bench_ab.log | cut -c60-190
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ab.log').read().strip().splitlines()[-1])
print({k:round(v['ms_per_step'],3) for k,v in d['families'].items() if 'wgrad' in k})
PY
done
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
