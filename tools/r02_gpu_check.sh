#!/bin/bash
# GPU box: the -m gpu suite, then the default bench line (both arms), outputs under gpurun_out/<tag>/
out=gpurun_out/${1:-r02w}; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/gpu_tests.log 2>&1; tail -3 $out/gpu_tests.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 1500 $out/bench.json; tail -3 $out/bench.err
