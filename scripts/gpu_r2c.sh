#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -12 gpurun_out/r2c_pytest.log
timeout 600 python scripts/fp32_accuracy_probe.py > gpurun_out/r2c_probe.log 2>&1; echo "probe rc=$?"
PN=2048 PK=512 timeout 600 python scripts/fp32_accuracy_probe.py >> gpurun_out/r2c_probe.log 2>&1
cat gpurun_out/r2c_probe.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2c_bench.err
