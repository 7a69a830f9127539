#!/bin/bash
# ncu captures on the current kernel sources: CTA-pair GEMM (A.[h|a], S^T.A), row-mode link loss, chained pooling
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_pool_chain.py -x -q -m gpu > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2r_bench.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['frac'], d['clocks']); print(d['roofline']['dominant_kernel'].get('chained_pooling')); print(d['roofline']['dominant_kernel']['tensor_contraction'])"
for w in ax2 tsa; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm2_kernel" --launch-skip 2 --launch-count 1 -o gpurun_out/prof_r2b_$w -f python scripts/gemm_probe.py $w 256 2 > gpurun_out/ncu_r2b_$w.log 2>&1; echo "ncu $w rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm2_kernel" --launch-skip 16 --launch-count 1 -o gpurun_out/prof_r2b_link -f python scripts/link_probe.py > gpurun_out/ncu_r2b_link.log 2>&1; echo "ncu link rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches_cfg4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r2b_launches.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/prof_r2b_* gpurun_out/r2b_launches_cfg4.csv
