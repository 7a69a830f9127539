#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/pk_dbg3.py > gpurun_out/dbg2.log 2>&1; echo rc=$?
grep -v Warn gpurun_out/dbg2.log | tail -60
