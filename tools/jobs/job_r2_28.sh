#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/debug_batch_wgrad.py 64 2 > gpurun_out/r2_28_dbg_64.log 2>&1; cat gpurun_out/r2_28_dbg_64.log | tail -4
timeout 300 python tools/debug_batch_wgrad.py 128 4 > gpurun_out/r2_28_dbg_128.log 2>&1; cat gpurun_out/r2_28_dbg_128.log | tail -4
