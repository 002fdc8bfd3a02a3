#!/bin/bash
# pytest -m gpu, then bench (ours + reference) at N=1, then launch list + ncu of the top kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest EXIT $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ours_n1.json 2> gpurun_out/bench_ours_n1.err; echo "bench ours EXIT $?"; tail -3 gpurun_out/bench_ours_n1.err; cat gpurun_out/bench_ours_n1.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "bench ref EXIT $?"; tail -3 gpurun_out/bench_ref_n1.err; cat gpurun_out/bench_ref_n1.json
