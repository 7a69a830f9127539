#!/bin/bash
mkdir -p gpurun_out
for g in on off; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --graph $g > gpurun_out/r2k_bench_$g.json 2> gpurun_out/r2k_bench_$g.err; echo "bench $g rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2k_bench_$g.json') if l.startswith('{')][-1]); print('graph $g:', d['ms_per_step'], d['config']['cuda_graph'], d['gpu_launches'], d['clocks'])"
tail -2 gpurun_out/r2k_bench_$g.err
done
