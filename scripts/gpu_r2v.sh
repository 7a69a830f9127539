#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2v_pytest.log
for c in ew8 ew4; do
if [ $c = ew4 ]; then export GP_TAIL_EW4=1; else unset GP_TAIL_EW4; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2v_bench_$c.json 2> gpurun_out/r2v_bench_$c.err; echo "bench $c rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2v_bench_$c.json') if l.startswith('{')][-1]); print('$c:', d['ms_per_step'], d['roofline']['frac'], d['clocks'])
for r in d['roofline']['kernels'][:18]: print('  ', r['entry'], r['shape'][:60], r['launches'], round(r['ms'],3), r.get('bound'), round(r.get('frac',0),3))"
done
