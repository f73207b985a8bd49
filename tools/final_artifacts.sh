#!/bin/bash
# Regenerates the measurement artefacts under gpurun_out/ on one B200 (copy what is to be judged into profiles/):
# bench lines (c2, reference arm, c3), c5 / c3 timings, rollout + covariance, single-series API timings, and the ncu
# launch lists of the c2 bench command and of one c5 evaluation.
set -x
timeout 200 python bench.py > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/f_bench_c2_reference.json 2>> gpurun_out/f_bench_c2.err
timeout 200 python bench.py --workload c3 --no-rollout > gpurun_out/f_bench_c3.json 2>> gpurun_out/f_bench_c2.err
timeout 200 python tools/bench_c5.py --cpu > gpurun_out/f_c5_c3_timing.json 2>&1
timeout 200 python tools/bench_rollout.py > gpurun_out/f_rollout_cov.json 2>&1
timeout 200 python tools/bench_single.py > gpurun_out/f_single_series.json 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/f_launches_c2_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-long > /dev/null 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/f_launches_c5_all.csv python tools/run_c5_once.py 8192 2 > /dev/null 2>&1
python tools/c5_launch_filter.py gpurun_out/f_launches_c5_all.csv gpurun_out/f_launches_c5_large.csv
ls -la gpurun_out/f_*
