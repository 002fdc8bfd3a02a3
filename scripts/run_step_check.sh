#!/bin/bash
# model-level checks + N=1 bench after model / op changes (one GPU)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "gpt2 or smoke or ddp_direct or layernorm or adamw" 2>&1 | tail -3
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ours_n1.json 2> gpurun_out/bench_ours_n1.err; cut -c1-220 gpurun_out/bench_ours_n1.json
timeout 200 python scripts/trace_step.py ours 2>&1 | grep -v Warning | sed -n 2,2p
