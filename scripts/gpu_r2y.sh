#!/bin/bash
timeout 300 python scripts/lbwd_probe.py 64 5000 1256 0; GP_LBWD_NOCFG=1 timeout 300 python scripts/lbwd_probe.py 64 5000 1256 0
timeout 300 python scripts/lbwd_probe.py 64 5000 1256 0 bf16; GP_LBWD_NOCFG=1 timeout 300 python scripts/lbwd_probe.py 64 5000 1256 0 bf16
timeout 600 python -m pytest tests/test_gpu_layer_bwd.py tests/test_gpu_baseline_shapes.py tests/test_gpu_fused_rows.py -q -m gpu 2>&1 | tail -2
