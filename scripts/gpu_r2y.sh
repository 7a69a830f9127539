#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_tc.py tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -q -m gpu 2>&1 | tail -2
for c in a b; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_$c.json 2> gpurun_out/r2y_bench_$c.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2y_bench_$c.json') if l.startswith('{')][-1]); print('$c:', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
for r in d['roofline']['kernels']:
  if 'bgemm_bf16x' in r['entry']: print('  ', r['shape'][:60], r['launches'], round(r['ms'],3))"
done
