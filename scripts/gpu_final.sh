#!/bin/bash
# end-of-round verification: what the driver runs (GPU tests, smoke, both bench arms), plus the launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
( time timeout 900 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err ) 2> gpurun_out/final_time_ref.txt; echo "ref rc=$?"; grep real gpurun_out/final_time_ref.txt; tail -c 600 gpurun_out/final_bench_ref.json
( time timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err ) 2> gpurun_out/final_time.txt; echo "bench rc=$?"; grep real gpurun_out/final_time.txt
python -c "
import json; d=json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
print('e2e', d['e2e']['value'], 'edge', d['e2e_edge_feed']['value'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
print('traffic', d['roofline']['traffic'], d['roofline']['dominant_kernel']['frac'], d['roofline']['dominant_kernel'].get('traffic_source'))
print('enz', d['enzymes_regime']['ms_per_step'], d['enzymes_regime']['frac_of_hbm_peak'])"
