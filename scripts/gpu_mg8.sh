#!/bin/bash
mkdir -p gpurun_out
N=${NG:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err; echo "bench rc=$?"
python - <<P
import json
txt=open('gpurun_out/mg${N}_bench.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('n_gpus', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'dp_check', d.get('dp_check'))
for k in ('e2e','e2e_u8_feed','e2e_edge_feed'):
    e=d.get(k) or {}
    print(k, {kk: e.get(kk) for kk in ('value','ms_per_step','strategy','error','host_pack_threads','host_pack_gbs','h2d_gbs')})
P
tail -3 gpurun_out/mg${N}_bench.err
