#!/bin/bash
# DDP gradient-proof run: bash scripts/run_ddp_verify.sh <N>   (ddp_debug, engines_check x3, exposed comm)
N=${1:-2}
mkdir -p gpurun_out
TR() { local t=$1; shift; local port=$1; shift; timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; echo "EXIT $?" >&2; }
F='Warning|warn|^$|\*\*\*|OMP_NUM'
TR 150 29518 scripts/ddp_debug.py > gpurun_out/ddp_debug_w$N.log 2>&1
grep -vE "$F" gpurun_out/ddp_debug_w$N.log | grep -E "^A raw|mismatching|DONE|Error" | tail -12 | cut -c1-300
for rep in 1 2 3; do
  TR 150 2951$rep scripts/engines_check.py > gpurun_out/engines_check_w${N}_rep$rep.log 2>&1
  grep -vE "$F" gpurun_out/engines_check_w${N}_rep$rep.log | grep -E "OK|FAIL|Error" | tr '\n' ';' | cut -c1-600; echo
done
TR 150 29519 scripts/exposed_comm.py > gpurun_out/exposed_comm_w$N.log 2>&1; grep -E "^\{|Error" gpurun_out/exposed_comm_w$N.log | tail -1 | cut -c1-400
