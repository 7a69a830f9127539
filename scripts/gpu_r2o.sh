#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/r2o_tc2.log 2>&1; echo "tc2 tests rc=$?"
tail -8 gpurun_out/r2o_tc2.log
timeout 300 python scripts/gemm_probe.py all 256 5 > gpurun_out/r2o_probe_pair.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2o_probe_pair.log
GP_NO_PAIR=1 timeout 300 python scripts/gemm_probe.py all 256 5 > gpurun_out/r2o_probe_nopair.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2o_probe_nopair.log
