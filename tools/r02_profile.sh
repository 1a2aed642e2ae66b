#!/bin/bash
# GPU box, round 2: fresh `ncu --set full` captures of the CURRENT kernels (the round-1 table described kernels that no longer exist)
# usage: bash tools/r02_profile.sh <tag> [tests]
TAG=${1:-r02a}; O=gpurun_out/$TAG; mkdir -p $O
if [ "$2" = "tests" ]; then timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $O/gpu_tests.log; fi
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:^k_cvp --launch-skip 1 --launch-count 1 -o $O/cvp_batch_1m python tools/prof_batch.py 1000 296 > $O/ncu_batch.log 2>&1; tail -2 $O/ncu_batch.log
timeout 600 $NCU -k regex:^k_cvp_grid --launch-skip 1 --launch-count 1 -o $O/cvp_grid_5m python tools/gpu_sweeps.py 2240 -1:1.8 > $O/ncu_grid.log 2>&1; tail -2 $O/ncu_grid.log
timeout 600 $NCU -k regex:"^k_layers|^k_inflate|^k_dijkstra_grid" --launch-skip 3 --launch-count 3 -o $O/misc_5m python tools/prof_misc.py 2240 > $O/ncu_misc.log 2>&1; tail -3 $O/ncu_misc.log
for f in $O/*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
ls -la $O
