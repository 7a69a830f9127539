#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_feed.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'])
for k in ('e2e','e2e_u8_feed','e2e_edge_feed'):
    e=d.get(k) or {}
    print(k, {kk: e.get(kk) for kk in ('value','ms_per_step','h2d_bytes_per_step','strategy','error')})
"
tail -3 gpurun_out/r2f_bench.err
