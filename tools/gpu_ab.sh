mkdir -p gpurun_out
for k in 3 1 2 0; do
GNAS_PW1=$k timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "pw1=$k rc=$?"; tail -1 gpurun_out/bench.log | cut -c60-190
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:(round(v['ms_per_step'],3), v['launches_per_step']) for k,v in d['families'].items() if 'pre_' in k})
PY
done
