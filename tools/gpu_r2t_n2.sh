#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time python -m pytest tests/test_parity_gpu.py -k "in_library" -q -x ) > $O/r2t_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2t_pytest.txt
python -c "import bench; bench.build_workload('config5')" > /dev/null 2>&1
( time SCDC_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-warmup 1 --no-cpu --no-extra ) > $O/r2t_bench_n2_t.json 2> $O/r2t_bench_n2_t.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-warmup 1 --no-cpu --no-extra ) > $O/r2t_bench_n2.json 2> $O/r2t_bench_n2.err
( time SCDC_TIMING=1 python bench.py --no-cpu --no-extra --steps 2 --warmup 1 --e2e-warmup 1 ) > $O/r2t_bench_n1.json 2> $O/r2t_bench_n1.err
grep -E "passed|failed|rc=" $O/r2t_pytest.txt | tail -3
