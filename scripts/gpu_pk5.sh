#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layer_fwd_kernel|layer_bwd_kernel|pool_fwd_kernel|pool_bwd_kernel" --launch-skip 14 --launch-count 14 -o gpurun_out/prof_pk5 -f python scripts/pk_profile.py > gpurun_out/pk5_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/pk5_ncu.log
ls -la gpurun_out/prof_pk5.ncu-rep
