#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layer_fwd_row_kernel" --launch-skip 5 --launch-count 2 -o gpurun_out/prof_pk8 -f python scripts/pk_profile.py > gpurun_out/pk8_ncu.log 2>&1; echo "ncu rc=$?"
