#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
for K1 in 512 1024; do
  echo "== K1_THREADS=$K1" >> $O/r2x_k1.txt
  SCDC_K1_THREADS=$K1 python bench.py --no-cpu --no-extra --no-e2e --steps 2 --warmup 1 >> $O/r2x_k1.txt 2>> $O/r2x_k1.err
done
( time python -m pytest tests -m gpu -q -x ) > $O/r2x_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2x_pytest.txt
python bench.py --no-cpu --no-extra --steps 2 --warmup 1 --e2e-warmup 1 > $O/r2x_bench.json 2> $O/r2x_bench.err
tail -3 $O/r2x_pytest.txt
