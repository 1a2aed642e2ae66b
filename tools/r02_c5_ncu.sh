#!/bin/bash
# GPU box: ncu capture (short metric list) of the three big kernels on the 50 M terrain
out=gpurun_out/${1:-r02c5d}; mkdir -p $out
timeout 700 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:'k_cvp_grid|k_layers_pf|k_dijkstra_grid' -c 3 --csv --log-file $out/c5_50m_metrics.csv python tools/gpu_config5_ncu.py 7072 > $out/ncu_run.log 2>&1
tail -3 $out/ncu_run.log; grep -c k_ $out/c5_50m_metrics.csv
