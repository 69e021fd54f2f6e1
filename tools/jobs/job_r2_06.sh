#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python tools/dgrad_fused_probe.py > gpurun_out/r2_06_probe.log 2>&1
cat gpurun_out/r2_06_probe.log
