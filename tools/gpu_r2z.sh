#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
B="python bench.py --no-cpu --no-extra --no-e2e --steps 2 --warmup 1"
for S in 0 1; do
  echo "== config5 STAGED=$S" >> $O/r2z_sweep.txt; SCDC_P1_STAGED=$S $B >> $O/r2z_sweep.txt 2>> $O/r2z_sweep.err
done
( time SCDC_P1_STAGED=1 python -m pytest tests -m gpu -q -x ) > $O/r2z_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2z_pytest.txt
for S in 0 1; do
  echo "== config3 STAGED=$S" >> $O/r2z_sweep.txt; SCDC_P1_STAGED=$S $B --workload config3 >> $O/r2z_sweep.txt 2>> $O/r2z_sweep.err
done
for S in 0 1; do
  echo "== config5 fp32 STAGED=$S" >> $O/r2z_sweep.txt; SCDC_P1_STAGED=$S $B --precision fp32_fast >> $O/r2z_sweep.txt 2>> $O/r2z_sweep.err
done
tail -3 $O/r2z_pytest.txt
