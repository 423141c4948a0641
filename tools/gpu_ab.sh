mkdir -p gpurun_out
for k in 511 503 495 511; do
GNAS_TC_PART=$k timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "tc_part=$k rc=$?"; tail -1 gpurun_out/bench.log | cut -c60-190
done
