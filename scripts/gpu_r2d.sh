#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2d.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r2d.log 2>&1; echo "ncu rc=$?"
