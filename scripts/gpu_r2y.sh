#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/lbwd_probe.py; GP_LBWD_NOFULL=1 timeout 300 python scripts/lbwd_probe.py
timeout 300 python scripts/lbwd_probe.py; GP_LBWD_NOFULL=1 timeout 300 python scripts/lbwd_probe.py
timeout 600 python -m pytest tests/test_gpu_layer_bwd.py tests/test_gpu_bf16.py -q -m gpu 2>&1 | tail -2
