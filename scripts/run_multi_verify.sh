#!/bin/bash
# multi-GPU verification after collective-kernel changes: bash scripts/run_multi_verify.sh <N> [bench-extra-args]
N=${1:-2}; shift
mkdir -p gpurun_out
TR() { local t=$1; shift; local port=$1; shift; timeout $t python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; echo "EXIT $?" >&2; }
F='Warning|warn|^$|\*\*\*|OMP_NUM'
if [ -z "$QUICK" ]; then
TR 150 29511 scripts/symm_check.py > gpurun_out/symm_check_w$N.log 2>&1
grep -vE "$F" gpurun_out/symm_check_w$N.log | grep -E "multicast|ALL_OK|FAIL|Error|gemm_rs|ag_gemm|'MiB': 25|'MiB': 256" | cut -c1-300 | tail -8
TR 150 29512 scripts/tp_check.py > gpurun_out/tp_check_w$N.log 2>&1
grep -vE "$F" gpurun_out/tp_check_w$N.log | grep -v "spin wait" | tail -4 | cut -c1-300
fi
TR 150 29513 scripts/engines_check.py > gpurun_out/engines_check_w$N.log 2>&1
grep -vE "$F" gpurun_out/engines_check_w$N.log | grep -E "OK|FAIL|Error" | tr '\n' ';' | cut -c1-900; echo
TR 400 29514 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_ours_n$N.json 2> gpurun_out/bench_ours_n$N.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_ours_n$N.json").read().strip().splitlines()[-1])
    keep={k:d.get(k) for k in ("value","ms_per_step","exposed_comm_ms","ms_per_step_without_collective","grad_check_rel","grad_check_rel_l2","params_identical_across_ranks","fused_reduce_optimizer","grad_check_error","exposed_comm_error","final_loss")}
    keep["e2e_ms"]=d.get("e2e",{}).get("ms_per_step")
    print(keep)
    if "other_configs" in d: print(json.dumps(d["other_configs"])[:2500])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_ours_n$N.err").read()[-1500:])
PY
