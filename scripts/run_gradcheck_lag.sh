#!/bin/bash
# fused reduce+optimizer step with the direct gradient route, host far ahead of the device (2 GPUs)
mkdir -p gpurun_out
TDP_FUSED_OPT=1 TDP_BENCH_GPU_LAG=0.3 TDP_BENCH_GRAD_DETAIL=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --other-configs off > gpurun_out/bench_gclag.json 2> gpurun_out/bench_gclag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_gclag.json").read().strip().splitlines()[-1])
err=open("gpurun_out/bench_gclag.err").read()
i=err.find("[grad detail rank 0] ")
per=None
if i>=0:
    arr,_=json.JSONDecoder().raw_decode(err[i+len("[grad detail rank 0] "):])
    per=[(x['bucket'], round(x['l2'],3), round(x['got_vs_local_l2'],3)) for x in arr]
print("fused+direct under lag:", {k:d.get(k) for k in ("grad_check_rel","grad_check_rel_l2","params_identical_across_ranks","fused_reduce_optimizer")}, per)
PY
