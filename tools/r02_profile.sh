#!/bin/bash
# GPU box, round 2: `ncu --set full` captures of the CURRENT kernels + the launch list of the bench command
# usage: bash tools/r02_profile.sh <tag>
TAG=${1:-r02z}; O=gpurun_out/$TAG; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:^k_cvp_batch --launch-skip 1 --launch-count 1 -o $O/cvp_batch_1m python tools/prof_batch.py 1000 592 > $O/ncu_batch.log 2>&1; tail -1 $O/ncu_batch.log
timeout 600 $NCU -k regex:^k_cvp_grid --launch-skip 1 --launch-count 1 -o $O/cvp_grid_5m python tools/gpu_sweeps.py 2240 -1:0 > $O/ncu_grid.log 2>&1; tail -1 $O/ncu_grid.log
timeout 600 $NCU -k regex:"^k_layers|^k_inflate|^k_dijkstra_grid" --launch-skip 3 --launch-count 3 -o $O/misc_5m python tools/prof_misc.py 2240 > $O/ncu_misc.log 2>&1; tail -1 $O/ncu_misc.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch-goals 592 --batch-steps 1 > $O/bench_under_ncu.log 2>&1
ls -la $O
