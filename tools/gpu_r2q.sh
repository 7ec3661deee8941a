#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
bash tools/gpu_sweep.sh r2q > /dev/null
SCDC_TIMING=1 python bench.py --no-cpu --no-extra --steps 1 --warmup 1 --e2e-warmup 1 > $O/r2q_timing.json 2> $O/r2q_timing.err
cat $O/sweep_r2q.txt
