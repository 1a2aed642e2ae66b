#!/bin/bash
# GPU box: BASELINE config 5 (50 M-vertex terrain) -- the bench line at --size 7072 (headline plan with its parity block against
# the oracle on the full 50 M map, Dijkstra, fused layers, inflation, dynamic cycle, 1M sub-mesh spot check), then one ncu
# --set full capture of the three big kernels at that size.
out=gpurun_out/${1:-r02c5}; mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $out/gpu.txt
( time timeout 1500 python bench.py --size 7072 --steps 3 --warmup 3 --batch-goals 0 --no-config3 > $out/bench_7072.json ) 2> $out/bench_7072.err
tail -c 600 $out/bench_7072.json
( time timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cvp_grid|k_layers_pf|k_dijkstra_grid' -c 3 -f -o $out/c5_50m python tools/gpu_config5_ncu.py 7072 > $out/ncu_run.log ) 2>> $out/ncu_run.log
tail -5 $out/ncu_run.log
ls -la $out
