#!/bin/bash
# A/B of the data-parallel step: bash scripts/run_ab_n.sh <N>
#   fused reduce+optimizer vs all-reduce + AdamW  x  collective CTA shape (128 / 512 threads)
N=${1:-2}
mkdir -p gpurun_out
i=0
for thr in 128 512; do for mode in 1 0; do
  i=$((i+1))
  TDP_COLL_THREADS=$thr TDP_FUSED_OPT=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --other-configs off > gpurun_out/bench_ab_n${N}_f${mode}_t$thr.json 2> gpurun_out/bench_ab_n${N}_f${mode}_t$thr.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_ab_n${N}_f${mode}_t$thr.json").read().strip().splitlines()[-1])
print("threads=$thr fused=$mode", {k:d.get(k) for k in ("ms_per_step","exposed_comm_ms","ms_per_step_without_collective","grad_check_rel")})
PY
done; done
