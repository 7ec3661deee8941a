#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q --durations=5 ) > $O/r2j_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2j_pytest.txt
bash tools/gpu_sweep.sh r2j > /dev/null
SCDC_TIMING=1 python bench.py --no-cpu --no-extra --steps 1 --warmup 1 --e2e-warmup 1 > $O/r2j_timing.json 2> $O/r2j_timing.err
tail -3 $O/r2j_pytest.txt
