#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
python -c "import bench; bench.build_workload('config5')" > /dev/null 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 3 --warmup 1 --e2e-warmup 1 --no-cpu --no-extra"
( SCDC_TIMING=1 SCDC_BCAST_NCCL=0 $T --master-port 29513 $B ) > $O/r2u_n2_ce.json 2> $O/r2u_n2_ce.err
( SCDC_TIMING=1 SCDC_BCAST_NCCL=1 $T --master-port 29514 $B ) > $O/r2u_n2_nccl.json 2> $O/r2u_n2_nccl.err
( SCDC_TIMING=1 SCDC_NO_NCCL=1 $T --master-port 29515 $B ) > $O/r2u_n2_nonccl.json 2> $O/r2u_n2_nonccl.err
