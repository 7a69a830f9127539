#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_tc.py tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py tests/test_gpu_fused_rows.py -q -m gpu 2>&1 | tail -2
for c in 0 1 0 1; do
if [ $c = 1 ]; then export GP_TAIL_EW4=1; else unset GP_TAIL_EW4; fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_t$c.json 2> gpurun_out/r2y_bench_t$c.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2y_bench_t$c.json') if l.startswith('{')][-1]); print('tail_ew4=$c:', d['ms_per_step'], d['roofline']['frac'], d['clocks'])
for r in d['roofline']['kernels']:
  if 'norm' in r['entry']: print('  ', r['entry'], r['shape'][:60], r['launches'], round(r['ms'],3), r.get('bound'), round(r.get('frac',0),3))"
done
