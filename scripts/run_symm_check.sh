#!/bin/bash
# usage: scripts/run_symm_check.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 200 python scripts/fused_check.py > gpurun_out/fused_check.log 2>&1; echo "fused EXIT $?"; tail -5 gpurun_out/fused_check.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/symm_check.py > gpurun_out/symm_check_w$N.log 2>&1; echo "symm EXIT $?"
tail -60 gpurun_out/symm_check_w$N.log
