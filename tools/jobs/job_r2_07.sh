#!/bin/bash
# round 2, job 7 (2 GPUs): PDL on by default, skip-add-only dgrad fusion; full GPU suite incl. the 2-rank NCCL test; benches N=1 and N=2
set -x
mkdir -p gpurun_out
PYT="python -u -m pytest -x -v --timeout 150 --timeout-method thread"
timeout 900 $PYT tests -m gpu > gpurun_out/r2_07_pytest.log 2>&1
tail -8 gpurun_out/r2_07_pytest.log
grep -E "worst relative|FAILED|Error" gpurun_out/r2_07_pytest.log | head
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_07_bench.log 2>&1
DG_PDL=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_07_bench_nopdl.log 2>&1
DG_DGRAD_BN_BWD=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_07_bench_nores.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_07_bench_n2.log 2>&1
DG_COMM=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-variants > gpurun_out/r2_07_bench_n2_torchdist.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload pix2pix_c4 --steps 8 --warmup 3 --no-cpu > gpurun_out/r2_07_bench_n2_pix2pix.log 2>&1
grep -h '"value"' gpurun_out/r2_07_bench*.log | cut -c1-220
