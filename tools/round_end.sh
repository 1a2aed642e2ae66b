#!/bin/bash
# GPU box: full parity suite, the bench line, the ncu launch list of the same command, full captures of the two grid kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; head -c 3000 gpurun_out/bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch-goals 296 --batch-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_cvp_grid --launch-skip 1 --launch-count 1 -f -o gpurun_out/cvp_grid_5m python tools/gpu_sweeps.py 2240 -1:1.8 > gpurun_out/ncu_cvp.log 2>&1; tail -2 gpurun_out/ncu_cvp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_dijkstra_grid --launch-skip 1 --launch-count 1 -f -o gpurun_out/dijkstra_grid_5m python tools/gpu_dij.py 2240 -1:3.0:-1 nocheck > gpurun_out/ncu_dij.log 2>&1; tail -2 gpurun_out/ncu_dij.log
