#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/r2t_bench_full.json 2> gpurun_out/r2t_bench_full.err ) 2> gpurun_out/r2t_time.txt; echo "bench rc=$?"; cat gpurun_out/r2t_time.txt
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2t_bench_full.json') if l.startswith('{')][-1])
print({k: d[k] for k in d if k not in ('roofline',)})
print(d['roofline']['frac'], d['roofline']['dominant_kernel'].get('traffic'), d['roofline']['dominant_kernel'].get('frac'))"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2t_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tc_gemm2_kernel<\(int\)256, \(int\)3" --launch-skip 14 --launch-count 1 -o gpurun_out/prof_r2b_link -f python scripts/link_probe.py > gpurun_out/ncu_r2b_link.log 2>&1; echo "ncu link rc=$?"
ls -la gpurun_out/prof_r2b_link.ncu-rep
