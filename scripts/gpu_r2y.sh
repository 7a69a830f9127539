#!/bin/bash
mkdir -p gpurun_out
for m in "" bf16; do
timeout 300 python scripts/lbwd_probe.py 256 2048 128 1 $m; GP_LBWD_NOCFG=1 timeout 300 python scripts/lbwd_probe.py 256 2048 128 1 $m
done
timeout 600 python -m pytest tests/test_gpu_layer_bwd.py tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py -q -m gpu 2>&1 | tail -2
for c in 0 1; do
if [ $c = 1 ]; then export GP_LBWD_NOCFG=1; else unset GP_LBWD_NOCFG; fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_c$c.json 2> gpurun_out/r2y_bench_c$c.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2y_bench_c$c.json') if l.startswith('{')][-1]); print('nocfg=$c:', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
for r in d['roofline']['kernels']:
  if 'layer_bwd' in r['entry']: print('  ', r['shape'][:60], r['launches'], round(r['ms'],3))"
done
