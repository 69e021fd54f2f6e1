#!/bin/bash
# round 2, job 69: captured frames are invalidated by new weights (test_video_graph_replay_matches_eager_frames)
mkdir -p gpurun_out
timeout 80 python -u -m pytest -x -q --timeout 70 --timeout-method thread tests/test_infer_gpu.py -k "graph_replay or shard_pipeline" > gpurun_out/r2_69_pytest.log 2>&1; tail -15 gpurun_out/r2_69_pytest.log
