#!/bin/bash
mkdir -p gpurun_out
T() { timeout "$@"; echo "EXIT $?"; }
T 250 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "not multi_gpu" > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
T 150 python scripts/fused_check.py > gpurun_out/fused_check.log 2>&1; tail -3 gpurun_out/fused_check.log
T 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/tp_check.py > gpurun_out/tp_check_w2.log 2>&1
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/tp_check_w2.log | tail -6
T 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/engines_check.py > gpurun_out/engines_check_w2.log 2>&1
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/engines_check_w2.log | tail -16
T 200 python scripts/profile_step.py ours > gpurun_out/profile_ours.log 2>&1; grep -A22 "^== ours" gpurun_out/profile_ours.log | cut -c1-150
T 200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ours_n1.json 2> gpurun_out/bench_ours_n1.err; tail -3 gpurun_out/bench_ours_n1.err; cat gpurun_out/bench_ours_n1.json
T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_ours_n2.json 2> gpurun_out/bench_ours_n2.err; tail -3 gpurun_out/bench_ours_n2.err; cat gpurun_out/bench_ours_n2.json
