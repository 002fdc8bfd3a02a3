#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/trace_step.py ours 2>&1 | grep -v Warning | tail -64
timeout 200 python scripts/trace_step.py reference 2>&1 | grep -v Warning | tail -64
