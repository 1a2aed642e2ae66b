#!/bin/bash
# GPU box: the bench line and the ncu launch list of the same command (optional: parity suite, A/B of the opt-in variants,
# full captures).  usage: bash tools/round_end.sh [tests] [ab] [full]
mkdir -p gpurun_out
for a in "$@"; do
if [ "$a" = "tests" ]; then timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/gpu_tests.log; fi
if [ "$a" = "ab" ]; then timeout 600 python tools/gpu_ab.py 2>&1 | tee gpurun_out/gpu_ab.log | tail -20; fi
done
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err; head -c 4500 gpurun_out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch-goals 296 --batch-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
for a in "$@"; do
if [ "$a" = "full" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_cvp_grid --launch-skip 1 --launch-count 1 -f -o gpurun_out/cvp_grid_5m python tools/gpu_sweeps.py 2240 -1:1.8 > gpurun_out/ncu_cvp.log 2>&1; tail -2 gpurun_out/ncu_cvp.log
fi
done
