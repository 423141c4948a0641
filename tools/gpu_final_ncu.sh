#!/bin/bash
# ncu evidence of the round: per-kernel counter CSVs of stand-alone units + one --set full report of the top kernels
mkdir -p gpurun_out
bash tools/ncu_units.sh "rhn" "cell 0" "cell 3" "cell 5"
python tools/prof_units.py rhn > gpurun_out/plain_rhn_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_rhn_bwd_gemm|k_rhn_fwd_tc|k_rhn_bwd_elem" -s 400 -c 6 -o gpurun_out/r02_rhn_full python tools/prof_units.py rhn > gpurun_out/ncu_rhn_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
