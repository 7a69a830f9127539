#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2y_pytest.log
for c in 0 1 0 1; do
if [ $c = 1 ]; then export GP_F32_GRADS=1; else unset GP_F32_GRADS; fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_g$c.json 2> gpurun_out/r2y_bench_g$c.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2y_bench_g$c.json') if l.startswith('{')][-1]); print('f32_grads=$c:', d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
grep -h "cfg4\|cfg3" gpurun_out/r2_parity_errors.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l)
    if r['precision']=='bf16': print(r['case'], 'ypred %.2e S %.2e grad %.3e cos %.5f'%(r['ypred'], r['S'] or 0, r['grad_flat'], r['grad_cos']))"
