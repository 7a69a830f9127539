#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/lbwd_probe.py 256 2048 128 0 bf16; timeout 300 python scripts/lbwd_probe.py 256 2048 512 0 bf16
timeout 900 python -m pytest tests/test_gpu_layer_bwd.py tests/test_gpu_fused_rows.py tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py tests/test_gpu_ops.py -q -m gpu 2>&1 | tail -2
GP_BENCH_MIN_SHARE=0.012 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_d.json 2> gpurun_out/r2y_bench_d.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2y_bench_d.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])
for r in d['roofline']['kernels']:
  if any(k in r['entry'] for k in ('bn_apply','softmax','layer_bwd')): print('  %-22s %-58s %2d %.3f %.3f'%(r['entry'][3:], r['shape'][:58], r['launches'], r['ms'], r.get('frac',0)))"
