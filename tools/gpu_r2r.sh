#!/bin/bash
# final round-2 validation on one GPU: GPU tests, smoke, default bench (config 5 headline + extras), reference arm, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q --durations=5 ) > $O/r2r_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2r_pytest.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2r_smoke.txt 2>&1
( time python bench.py --impl reference ) > $O/r2r_ref.json 2> $O/r2r_ref.err
( time python bench.py ) > $O/r2r_bench.json 2> $O/r2r_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2r_launches_config5.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2r_ncu0.log 2>&1
SCDC_TIMING=1 python bench.py --no-cpu --no-extra --steps 1 --warmup 1 --e2e-warmup 1 > $O/r2r_timing.json 2> $O/r2r_timing.err
tail -3 $O/r2r_pytest.txt; tail -2 $O/r2r_smoke.txt
