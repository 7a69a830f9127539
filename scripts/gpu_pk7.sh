#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/pk_profile.py > gpurun_out/pk7_profile.log 2>&1; echo "rc=$?"
tail -36 gpurun_out/pk7_profile.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pk7_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pk7_pytest.log
