#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layer_bwd_bn_cta2_kernel" --launch-skip 3 --launch-count 1 -o gpurun_out/prof_r2f_lbwd -f python scripts/lbwd_probe.py 256 2048 128 1 bf16 > gpurun_out/ncu_r2f_lbwd.log 2>&1; echo "ncu rc=$?"
GP_BENCH_NO_ENZ=1 GP_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --graph off > gpurun_out/ncu_r2f_launches.log 2>&1; echo "launch list rc=$?"
