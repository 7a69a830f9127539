#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2p_pytest.log
for c in pair nopair; do
if [ $c = nopair ]; then export GP_NO_PAIR=1; else unset GP_NO_PAIR; fi
GP_NO_CHAIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_bench_$c.json 2> gpurun_out/r2p_bench_$c.err; echo "bench $c rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2p_bench_$c.json') if l.startswith('{')][-1]); print('$c:', d['ms_per_step'], d['roofline']['frac'], d['clocks'])
for r in d['roofline']['kernels'][:16]: print('  ', r['entry'], r['shape'][:60], r['launches'], round(r['ms'],3), r.get('bound'), round(r.get('frac',0),3))"
done
