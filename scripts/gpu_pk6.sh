#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/packed_debug.py > gpurun_out/pk6_debug.log 2>&1; echo "debug rc=$?"
grep -E "worst|<<<|Error|error" gpurun_out/pk6_debug.log | tail -12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layer_fwd_row_kernel|layer_bwd_row_kernel" --launch-skip 10 --launch-count 10 -o gpurun_out/prof_pk6 -f python scripts/pk_profile.py > gpurun_out/pk6_ncu.log 2>&1; echo "ncu rc=$?"
