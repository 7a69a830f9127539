#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_bf16_small_errors.jsonl gpurun_out/r2_parity_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2g_pytest.log
timeout 900 python scripts/cfg1_batch_sweep.py --batches 20,256,1024,4096,16384 > gpurun_out/r2g_sweep.md 2> gpurun_out/r2g_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r2g_sweep.md | tail -8
GP_NO_PACKED=1 timeout 900 python scripts/cfg1_batch_sweep.py --batches 20,4096,16384 > gpurun_out/r2g_sweep_dense.md 2>/dev/null; tail -4 gpurun_out/r2g_sweep_dense.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kernel" --launch-skip 400 -c 120 --csv --log-file gpurun_out/launches_r2g_cfg1_b4096.csv python scripts/cfg1_batch_sweep.py --batches 4096 --steps 2 > gpurun_out/ncu_r2g.log 2>&1; echo "ncu rc=$?"
