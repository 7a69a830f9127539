#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity_errors.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pk2_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pk2_pytest.log
timeout 600 python scripts/cfg1_batch_sweep.py --batches 20,256,4096,16384 > gpurun_out/pk2_sweep.md 2> gpurun_out/pk2_sweep.err; echo "sweep rc=$?"
cat gpurun_out/pk2_sweep.md; tail -5 gpurun_out/pk2_sweep.err
GP_NO_PACKED=1 timeout 600 python scripts/cfg1_batch_sweep.py --batches 4096 > gpurun_out/pk2_sweep_dense.md 2>&1
tail -3 gpurun_out/pk2_sweep_dense.md
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gp|pk" -c 200 --csv --log-file gpurun_out/launches_pk2.csv python scripts/cfg1_batch_sweep.py --batches 4096 --steps 1 > gpurun_out/ncu_pk2.log 2>&1; echo "ncu rc=$?"
