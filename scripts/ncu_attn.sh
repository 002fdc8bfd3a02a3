#!/bin/bash
# one ncu --set full capture of each attention kernel (fwd, bwd dQ, bwd dK/dV) + text summaries
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_ -s 3 -c 3 -f -o gpurun_out/prof_attn python scripts/attn_prof.py > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
