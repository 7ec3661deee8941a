#!/bin/bash
# 2-GPU box: the in-library multi-GPU call (tests) and the N = 2 bench line (torchrun)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r2g_smi.txt
( time NCCL_DEBUG=INFO python -m pytest tests/test_parity_gpu.py -k "in_library" -q -x ) > $O/r2g_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2g_pytest.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 ) > $O/r2g_bench_n2.json 2> $O/r2g_bench_n2.err
( time SCDC_NO_NCCL=1 SCDC_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-warmup 1 ) > $O/r2g_bench_n2_nonccl.json 2> $O/r2g_bench_n2_nonccl.err
grep -E "passed|failed|rc=" $O/r2g_pytest.txt | tail -3
