#!/bin/bash
mkdir -p gpurun_out
for w in cfg1_enzymes_like cfg2_dd_base cfg3_dd_diffpool_p2 cfg5_ragged_64x5000; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench_$w.json 2> gpurun_out/r2w_bench_$w.err; echo "bench $w rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2w_bench_$w.json') if l.startswith('{')][-1]); print('$w:', d['ms_per_step'], d['value'], d['config'].get('cuda_graph'), d['roofline']['frac'], (d.get('e2e') or {}).get('value'))"
tail -2 gpurun_out/r2w_bench_$w.err
done
