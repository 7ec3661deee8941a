#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time SCDC_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-warmup 1 ) > $O/r2l_bench_n2.json 2> $O/r2l_bench_n2.err
bash tools/gpu_sweep.sh r2l > /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2l_launches_config5.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2l_ncu0.log 2>&1
cat $O/sweep_r2l.txt
