#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r2v_smi.txt
python -c "import bench; bench.build_workload('config5')" > /dev/null 2>&1   # build the cached problem once, not 8 times
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time SCDC_TIMING=1 $T --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 1 --e2e-warmup 1 --no-cpu --no-extra ) > $O/r2v_bench_n8.json 2> $O/r2v_bench_n8.err
( time $T --master-port 29522 bench.py --gpus 8 --steps 1 --warmup 1 --e2e-warmup 1 --no-cpu --no-extra --precision fp32_fast ) > $O/r2v_bench_n8_fp32.json 2> $O/r2v_bench_n8_fp32.err
tail -2 $O/r2v_bench_n8.err
