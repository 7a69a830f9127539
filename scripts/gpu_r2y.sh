#!/bin/bash
mkdir -p gpurun_out
for m in "" bf16; do
timeout 300 python scripts/lbwd_probe.py 256 2048 512 0 $m; GP_LBWD_NOCFG=1 timeout 300 python scripts/lbwd_probe.py 256 2048 512 0 $m
timeout 300 python scripts/lbwd_probe.py 256 2048 128 0 $m; GP_LBWD_NOCFG=1 timeout 300 python scripts/lbwd_probe.py 256 2048 128 0 $m
done
timeout 600 python -m pytest tests/test_gpu_layer_bwd.py tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py tests/test_gpu_fused_rows.py -q -m gpu 2>&1 | tail -2
