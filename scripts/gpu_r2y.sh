#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2y_pytest.log
for w in cfg5_ragged_64x5000 cfg4_diffpool_256x2048; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2y_bench_$w.json 2> gpurun_out/r2y_bench_$w.err; echo "bench $w rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2y_bench_$w.json') if l.startswith('{')][-1]); print('$w:', d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
