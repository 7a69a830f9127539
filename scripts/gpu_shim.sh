#!/bin/bash
mkdir -p gpurun_out
cd /root/repo
( time timeout 900 python -m graph_pooling_b200.shim --reference .ref_scratch --seed 0 -- --bmname=ENZYMES --datadir=.ref_scratch/data --method=soft-assign --max-nodes=100 --num-classes=6 --hidden-dim=30 --output-dim=30 --assign-ratio=0.1 --num-pool=1 --linkpred --epochs=5 --num_workers=1 ) > gpurun_out/r2_shim_train.log 2>&1; echo "shim rc=$?"
grep -E "epoch time|Validation  accuracy|real" gpurun_out/r2_shim_train.log | tail -8
grep -c "Epoch:" gpurun_out/r2_shim_train.log
( time GP_NO_PACKED=1 timeout 900 python -m graph_pooling_b200.shim --reference .ref_scratch --seed 0 -- --bmname=ENZYMES --datadir=.ref_scratch/data --method=soft-assign --max-nodes=100 --num-classes=6 --hidden-dim=30 --output-dim=30 --assign-ratio=0.1 --num-pool=1 --linkpred --epochs=5 --num_workers=1 ) > gpurun_out/r2_shim_train_dense.log 2>&1; echo "shim dense rc=$?"
grep -E "epoch time|real" gpurun_out/r2_shim_train_dense.log | tail -4
