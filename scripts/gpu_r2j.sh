#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_feed.py tests/test_gpu_bf16.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2j_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2j_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'])
for k in ('e2e','e2e_edge_feed'):
    e=d.get(k) or {}
    print(k, {kk: e.get(kk) for kk in ('value','ms_per_step','h2d_bytes_per_step','edge_id_bytes','feature_dtype','error')})
"
tail -3 gpurun_out/r2j_bench.err
