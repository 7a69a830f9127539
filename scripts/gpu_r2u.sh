#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_loss_options.py tests/test_gpu_bf16.py -x -q -m gpu > gpurun_out/r2u_tc2.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2u_tc2.log
timeout 300 python scripts/link_probe.py > gpurun_out/r2u_link.log 2>&1; cat gpurun_out/r2u_link.log
GP_LINK_STAGES4=1 timeout 300 python scripts/link_probe.py > gpurun_out/r2u_link4.log 2>&1; cat gpurun_out/r2u_link4.log
