#!/bin/bash
mkdir -p gpurun_out
GP_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2h_cfg1_b4096.csv python scripts/cfg1_batch_sweep.py --batches 4096 --steps 1 > gpurun_out/ncu_r2h.log 2>&1; echo "ncu list rc=$?"
GP_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"layer_fwd_row_kernel|layer_bwd_kernel|pool_fwd_kernel|pool_bwd_kernel|layer_fwd_kernel" -c 14 -o gpurun_out/prof_r2h_packed -f python scripts/cfg1_batch_sweep.py --batches 4096 --steps 1 > gpurun_out/ncu_r2h_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_r2h_packed.ncu-rep
