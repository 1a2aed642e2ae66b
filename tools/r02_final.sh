#!/bin/bash
# GPU box: the -m gpu suite, smoke(), the default bench line of both arms, a launch list of one bench step.
out=gpurun_out/${1:-r02final}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/gpu_tests.log 2>&1; tail -2 $out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json; tail -2 $out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 300 $out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --batch-goals 0 --no-config3 --no-cpu-baseline > $out/ncu_bench.log 2>&1
grep -c k_ $out/launches.csv
