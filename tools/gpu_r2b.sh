#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -q --durations=8 -x ) > $O/r2b_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2b_pytest.txt
( time python bench.py --no-cpu --steps 3 --warmup 2 ) > $O/r2b_bench.json 2> $O/r2b_bench.err
tail -5 $O/r2b_pytest.txt
