#!/bin/bash
# GPU box: BASELINE config 5 (50 M-vertex terrain) -- the bench line at --size 7072 (headline plan with its parity block against
# the oracle on the full 50 M map, Dijkstra, fused layers, inflation, dynamic cycle, ray casting, 1M sub-mesh spot check), then an
# ncu capture of the three big kernels at that size (a short metric list: under --set full the cooperative wavefront kernels'
# counters come back as NaN at this size).
out=gpurun_out/${1:-r02c5}; mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $out/gpu.txt
( time timeout 1500 python bench.py --size 7072 --steps 3 --warmup 3 --batch-goals 0 --no-config3 > $out/bench_7072.json ) 2> $out/bench_7072.err
tail -c 600 $out/bench_7072.json
[ "$2" = "noncu" ] && exit 0
( time timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:'k_cvp_grid|k_layers_pf|k_dijkstra_grid' -c 3 --csv --log-file $out/c5_50m_metrics.csv python tools/gpu_config5_ncu.py 7072 > $out/ncu_run.log ) 2>> $out/ncu_run.log
tail -5 $out/ncu_run.log; tail -30 $out/c5_50m_metrics.csv | cut -c1-250
