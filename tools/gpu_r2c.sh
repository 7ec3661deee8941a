#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
( time python -m pytest tests -m gpu -q --durations=8 ) > $O/r2c_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2c_pytest.txt
( time python bench.py --no-cpu --steps 2 --warmup 1 ) > $O/r2c_bench.json 2> $O/r2c_bench.err
SCDC_TIMING=1 python bench.py --no-cpu --no-extra --steps 1 --warmup 1 --e2e-warmup 1 > $O/r2c_timing.json 2> $O/r2c_timing.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2c_launches_config5.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2c_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2c_launches_config5_fp32.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e --precision fp32_fast > $O/r2c_ncu2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2c_launches_config3.csv \
  python bench.py --workload config3 --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2c_ncu3.log 2>&1
tail -5 $O/r2c_pytest.txt
