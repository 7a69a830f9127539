#!/bin/bash
mkdir -p gpurun_out
for w in 64 96 128 192; do
  echo "WINDOW=$w"; GP_PK_WINDOW=$w timeout 300 python scripts/cfg1_batch_sweep.py --batches 4096 2>/dev/null | tail -1
done
for pr in 32 128 256; do
  echo "POST_ROWS=$pr"; GP_PK_POST_ROWS=$pr timeout 300 python scripts/cfg1_batch_sweep.py --batches 4096 2>/dev/null | tail -1
done
