#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q > gpurun_out/final_mg_pytest.log 2>&1; echo "dp pytest rc=$?"; tail -2 gpurun_out/final_mg_pytest.log
for N in 2 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_mg${N}_bench.json 2> gpurun_out/final_mg${N}_bench.err; echo "bench $N rc=$?"
python - <<P
import json
txt=open('gpurun_out/final_mg${N}_bench.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('n_gpus', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'dp_check', d.get('dp_check'))
for k in ('e2e','e2e_u8_feed','e2e_edge_feed'):
    e=d.get(k) or {}
    print(k, {kk: e.get(kk) for kk in ('value','ms_per_step')})
P
done
