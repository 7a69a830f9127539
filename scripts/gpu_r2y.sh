#!/bin/bash
mkdir -p gpurun_out
for c in 0 1 0; do
if [ $c = 1 ]; then export GP_LBWD_NOCFG=1; else unset GP_LBWD_NOCFG; fi
GP_BENCH_MIN_SHARE=0.012 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_c$c.json 2> gpurun_out/r2y_bench_c$c.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2y_bench_c$c.json') if l.startswith('{')][-1]); print('nocfg=$c:', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2y_bench_c0.json') if l.startswith('{')][-1])
for r in d['roofline']['kernels']: print('  %-22s %-64s %2d %.3f %s %.3f'%(r['entry'][3:], r['shape'][:64], r['launches'], r['ms'], r.get('bound'), r.get('frac',0)))"
