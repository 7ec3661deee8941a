#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q -x ) > $O/r2z2_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2z2_pytest.txt
python bench.py --no-cpu --no-extra --steps 2 --warmup 1 --e2e-warmup 1 --workload config3 > $O/r2z2_config3.json 2> $O/r2z2_config3.err
tail -3 $O/r2z2_pytest.txt
