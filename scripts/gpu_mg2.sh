#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > gpurun_out/mg2_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/mg2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/mg2_bench.json 2> gpurun_out/mg2_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/mg2_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'dp_check', d.get('dp_check'))
for k in ('e2e','e2e_u8_feed','e2e_edge_feed'):
    e=d.get(k) or {}
    print(k, {kk: e.get(kk) for kk in ('value','ms_per_step','strategy','error')})
"
tail -3 gpurun_out/mg2_bench.err
