#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q --durations=5 ) > $O/r2f_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2f_pytest.txt
bash tools/gpu_sweep.sh r2f > /dev/null
tail -3 $O/r2f_pytest.txt
