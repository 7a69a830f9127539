#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --graph on > gpurun_out/r2e_bench_graph.json 2> gpurun_out/r2e_bench_graph.err; echo "bench graph rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_graph.json')); print('graph on:', d['ms_per_step'], d['config']['cuda_graph'], d['gpu_launches'], d.get('enzymes_regime'))"
tail -3 gpurun_out/r2e_bench_graph.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --graph off > gpurun_out/r2e_bench_off.json 2> gpurun_out/r2e_bench_off.err; echo "bench off rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_off.json')); print('graph off:', d['ms_per_step'], d['gpu_launches'])"
GP_BENCH_NO_ENZ=1 GP_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2e.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --graph off > gpurun_out/ncu_r2e.log 2>&1; echo "ncu rc=$?"
