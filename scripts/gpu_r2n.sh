#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_probe.py chain 256 5 > gpurun_out/r2n_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2n_probe.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2n_pytest.log
for c in chain nochain; do
if [ $c = nochain ]; then export GP_NO_CHAIN=1; else unset GP_NO_CHAIN; fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2n_bench_$c.json 2> gpurun_out/r2n_bench_$c.err; echo "bench $c rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2n_bench_$c.json') if l.startswith('{')][-1]); print('$c:', d['ms_per_step'], d['roofline']['frac'], d['clocks'])
for r in d['roofline']['kernels'][:14]: print('  ', r['entry'], r['shape'][:60], r['launches'], round(r['ms'],3), r.get('bound'), round(r.get('frac',0),3))"
done
unset GP_NO_CHAIN
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pool_chain_kernel" --launch-skip 4 --launch-count 1 -o gpurun_out/prof_r2_chain -f python scripts/gemm_probe.py chain 256 2 > gpurun_out/ncu_r2_chain.log 2>&1; echo "ncu chain rc=$?"
ls -la gpurun_out/prof_r2_chain.ncu-rep
