#!/bin/bash
# gradient-check A/B under device lag (2 GPUs): which ingredient breaks the eager fused step?
N=2
mkdir -p gpurun_out
run() {  # name, env...
  local name=$1; shift
  env "$@" TDP_BENCH_GPU_LAG=0.3 TDP_BENCH_GRAD_DETAIL=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --other-configs off ${EXTRA} > gpurun_out/bench_gcab_$name.json 2> gpurun_out/bench_gcab_$name.err
  python - <<PY
import json,re
d=json.loads(open("gpurun_out/bench_gcab_$name.json").read().strip().splitlines()[-1])
err=open("gpurun_out/bench_gcab_$name.err").read()
m=re.search(r'\[grad detail rank 0\] (\[.*?\])', err)
per=[(x['bucket'], round(x['l2'],3), round(x['got_vs_local_l2'],3)) for x in json.loads(m.group(1))] if m else None
print("$name", {k:d.get(k) for k in ("grad_check_rel","grad_check_rel_l2")}, per)
PY
}
run fused_nodirect TDP_FUSED_OPT=1 TDP_FUSED_WGRAD=0
EXTRA=--no-graph run plain_eager TDP_FUSED_OPT=0
EXTRA=--no-graph run plain_eager_nodirect TDP_FUSED_OPT=0 TDP_FUSED_WGRAD=0
