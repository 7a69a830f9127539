#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/pk_profile.py > gpurun_out/pk3_profile.log 2>&1; echo "rc=$?"
cat gpurun_out/pk3_profile.log | tail -60
