#!/bin/bash
mkdir -p gpurun_out
GP_BENCH_NO_ENZ=1 GP_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --graph off > gpurun_out/ncu_r2b_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_gemm2_kernel<256, 3, 1, 16" --launch-skip 14 --launch-count 1 -o gpurun_out/prof_r2b_link -f python scripts/link_probe.py > gpurun_out/ncu_r2b_link.log 2>&1; echo "ncu link rc=$?"
ls -la gpurun_out/prof_r2b_link.ncu-rep gpurun_out/r2b_launches_cfg4.csv
