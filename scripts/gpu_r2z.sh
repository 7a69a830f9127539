#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2z_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2z_bench.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
for r in d['roofline']['kernels']: print('  ', r['entry'], r['shape'][:70], r['launches'], round(r['ms'],3), r.get('bound'), round(r.get('frac',0),3))"
