#!/bin/bash
# last call of the round (one GPU): the whole GPU test-suite + a short headline run
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_ours_n1_final.json 2> gpurun_out/bench_ours_n1_final.err; cut -c1-200 gpurun_out/bench_ours_n1_final.json
