#!/bin/bash
# scratch A/B template for one GPU call: parity tests, then the bench (edit the environment switches per experiment)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for k in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -1 gpurun_out/bench.log | cut -c60-190
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['families'].items() if 'mix' in k or 'reduce' in k})
PY
done
