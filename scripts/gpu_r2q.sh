#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc2.py -x -q -m gpu > gpurun_out/r2q_tc2.log 2>&1; echo "tc2 tests rc=$?"
tail -8 gpurun_out/r2q_tc2.log
timeout 300 python scripts/link_probe.py > gpurun_out/r2q_link.log 2>&1; cat gpurun_out/r2q_link.log
GP_NO_UPPER_G=1 timeout 300 python scripts/link_probe.py > gpurun_out/r2q_link_full.log 2>&1; cat gpurun_out/r2q_link_full.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2q_pytest.log
GP_NO_CHAIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2q_bench.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['frac'], d['clocks'])
for r in d['roofline']['kernels'][:16]: print('  ', r['entry'], r['shape'][:60], r['launches'], round(r['ms'],3), r.get('bound'), round(r.get('frac',0),3))"
