#!/bin/bash
mkdir -p gpurun_out
( timeout 40 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "direct_weight or layernorm" 2>&1 | tail -4 ) &
timeout 50 python bench.py --gpus 1 --steps 6 --warmup 3 > gpurun_out/bench_ours_n1_fw.json 2> gpurun_out/bench_ours_n1_fw.err; tail -2 gpurun_out/bench_ours_n1_fw.err | cut -c1-300; cut -c1-420 gpurun_out/bench_ours_n1_fw.json
wait
