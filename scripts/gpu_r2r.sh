#!/bin/bash
# FINAL ncu captures on the final kernel sources: CTA-pair GEMM (A.[h|a], S^T.A), row-epilogue link loss, launch list
mkdir -p gpurun_out
for w in ax2 tsa; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm2_kernel" --launch-skip 2 --launch-count 1 -o gpurun_out/prof_r2f_$w -f python scripts/gemm_probe.py $w 256 2 > gpurun_out/ncu_r2f_$w.log 2>&1; echo "ncu $w rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_gemm2_kernel<\(int\)256, \(int\)3, \(int\)3" --launch-skip 14 --launch-count 1 -o gpurun_out/prof_r2f_link -f python scripts/link_probe.py > gpurun_out/ncu_r2f_link.log 2>&1; echo "ncu link rc=$?"
GP_BENCH_NO_ENZ=1 GP_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --graph off > gpurun_out/ncu_r2f_launches.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/prof_r2f_* gpurun_out/r2f_launches_cfg4.csv
