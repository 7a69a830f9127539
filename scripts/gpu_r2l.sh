#!/bin/bash
mkdir -p gpurun_out
for w in ax2 tsa; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm2_kernel" --launch-skip 2 --launch-count 1 -o gpurun_out/prof_r2_$w -f python scripts/gemm_probe.py $w 256 2 > gpurun_out/ncu_r2_$w.log 2>&1; echo "ncu $w rc=$?"
done
ls -la gpurun_out/prof_r2_*.ncu-rep
