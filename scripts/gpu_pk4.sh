#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/packed_debug.py > gpurun_out/pk4_debug.log 2>&1; echo "debug rc=$?"
grep -E "case|worst|<<<|Error|error" gpurun_out/pk4_debug.log | tail -30
timeout 600 python scripts/pk_profile.py > gpurun_out/pk4_profile.log 2>&1; echo "rc=$?"
tail -40 gpurun_out/pk4_profile.log
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_model.py tests/test_gpu_graphed.py tests/test_gpu_enzymes.py -m gpu -q -x > gpurun_out/pk4_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pk4_pytest.log
