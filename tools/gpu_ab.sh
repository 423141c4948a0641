mkdir -p gpurun_out
for m in 2 1 0; do
CHECK=0 GNAS_RHN_FUSED=$m timeout 300 python tools/rhn_bench.py > gpurun_out/r2_rhn_fused$m.txt 2>&1; echo "mode $m"; head -5 gpurun_out/r2_rhn_fused$m.txt | cut -c1-120
GNAS_RHN_FUSED=$m timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/bench_fused$m.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_fused$m.log | cut -c1-200; tail -3 gpurun_out/bench.err
done
GNAS_RHN_FUSED=2 timeout 600 python -m pytest tests/test_rnn.py tests/test_model.py -x -q -m gpu 2>&1 | tail -2
