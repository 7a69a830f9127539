#!/bin/bash
# chained S^T A S kernel: unit tests first (own timeout), then the suite and a bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pool_chain.py -x -q -m gpu > gpurun_out/r2m_chain.log 2>&1; echo "chain tests rc=$?"
tail -15 gpurun_out/r2m_chain.log
