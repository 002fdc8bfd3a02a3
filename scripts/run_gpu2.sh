#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python scripts/gemm_check.py > gpurun_out/gemm_check.log 2>&1; echo "gemm EXIT $?"; grep -E "ALL_OK|Error|error" gpurun_out/gemm_check.log | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/symm_check.py > gpurun_out/symm_check_w$N.log 2>&1; echo "symm EXIT $?"
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/symm_check_w$N.log | tail -40
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/tp_check.py > gpurun_out/tp_check_w$N.log 2>&1; echo "tp EXIT $?"
grep -vE "Warning|warn|^$|\*\*\*|OMP_NUM" gpurun_out/tp_check_w$N.log | tail -30
