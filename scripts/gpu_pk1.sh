#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/packed_debug.py > gpurun_out/pk1_debug.log 2>&1; echo "debug rc=$?"
tail -150 gpurun_out/pk1_debug.log
