#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q --durations=5 ) > $O/r2n_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2n_pytest.txt
( time python bench.py ) > $O/r2n_bench.json 2> $O/r2n_bench.err
SCDC_TIMING=1 python bench.py --no-cpu --no-extra --steps 1 --warmup 1 --e2e-warmup 1 > $O/r2n_timing.json 2> $O/r2n_timing.err
SCDC_NO_EARLY_PASS1=1 python bench.py --no-cpu --no-extra --steps 1 --warmup 1 --e2e-warmup 1 > $O/r2n_noearlyp1.json 2> $O/r2n_noearlyp1.err
tail -3 $O/r2n_pytest.txt
