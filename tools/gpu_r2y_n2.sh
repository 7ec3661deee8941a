#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time python -m pytest tests/test_parity_gpu.py tests/test_fp32_fast_gpu.py -k "in_library or multi_gpu or n_gpus" -q ) > $O/r2y_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2y_pytest.txt
tail -4 $O/r2y_pytest.txt
