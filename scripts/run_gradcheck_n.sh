#!/bin/bash
# short headline run with per-bucket gradient-check details: bash scripts/run_gradcheck_n.sh <N> [lag seconds]
N=${1:-8}
mkdir -p gpurun_out
for fused in 1 0; do
TDP_FUSED_OPT=$fused TDP_BENCH_GPU_LAG=${2:-0.3} TDP_BENCH_GRAD_DETAIL=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$fused bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --other-configs off > gpurun_out/bench_gc_n${N}_f$fused.json 2> gpurun_out/bench_gc_n${N}_f$fused.err
grep "grad detail rank 0\]" gpurun_out/bench_gc_n${N}_f$fused.err | python -c "
import sys,json,re
for ln in sys.stdin:
    m=re.search(r'\[grad detail rank (\d+)\] (\[.*?\])(?=\[grad detail|\$)', ln)
    if not m: print(ln[:400]); continue
    for d in json.loads(m.group(2)): print('fused=$fused rank',m.group(1),'bucket',d['bucket'],'rel %.3g l2 %.3g got_vs_local %.3g'%(d['rel'],d['l2'],d['got_vs_local_l2']))
"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_gc_n${N}_f$fused.json").read().strip().splitlines()[-1])
print("fused=$fused", {k:d.get(k) for k in ("ms_per_step","grad_check_rel","grad_check_rel_l2","params_identical_across_ranks")})
PY
done
