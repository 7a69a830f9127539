#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_baseline_shapes.py tests/test_gpu_tc2.py tests/test_gpu_loss_options.py tests/test_gpu_edge_feed.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2i_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench.json')); print('fused softmax:', d['ms_per_step'], d['roofline']['frac'])
for r in d['roofline']['kernels'][:14]: print('  ', r['entry'], r['shape'][:50], r['launches'], round(r['ms'],3))"
GP_NO_FUSED_SOFTMAX=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2i_bench_nofuse.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_nofuse.json')); print('unfused:', d['ms_per_step'])"
