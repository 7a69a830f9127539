#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity_errors.jsonl
timeout 600 python scripts/parity_table.py > gpurun_out/r2b_parity_table.md 2>&1; echo "table rc=$?"
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
