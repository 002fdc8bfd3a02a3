#!/bin/bash
# compute-sanitizer passes over the native kernels (one GPU; run on a B200 box:
#   gpurun --timeout 900 -- 'bash scripts/run_sanitizer.sh').
# memcheck: out-of-bounds / misaligned accesses (incl. TMA boxes clipped at ragged edges);
# racecheck: shared-memory hazards between the epilogue warps / staging buffers;
# synccheck: invalid barrier usage (named barriers, mbarrier, cluster barrier).
# (initcheck: TOOLS="memcheck racecheck synccheck initcheck")
# The kernels are big and the tool serialises them: keep the problem sizes small.
mkdir -p gpurun_out
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
SEL=${SEL:-"fused_epilogues or layernorm or adamw or (native_attention_forward and True-128)"}
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool"
  timeout ${LIMIT:-170} $SAN --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$SEL" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Race reported|Invalid|Uninitialized|hazard" gpurun_out/sanitizer_$tool.log | tail -5
done
