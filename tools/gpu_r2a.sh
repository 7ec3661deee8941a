#!/bin/bash
# round-2 baseline: GPU tests, default bench (config5 headline), reference arm, ncu launch list + full captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_smi.txt
( time python -m pytest tests -m gpu -x -q --durations=15 ) > $O/r2a_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2a_pytest.txt
( time python bench.py ) > $O/r2a_bench.json 2> $O/r2a_bench.err
( time python bench.py --impl reference ) > $O/r2a_ref.json 2> $O/r2a_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2a_launches_config5.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2a_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -c 1 -o $O/r2a_k_scatter_config5 -f \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2a_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_score_diff -c 1 -o $O/r2a_k_score_diff_config5 -f \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2a_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -c 1 -o $O/r2a_k_scatter_config3 -f \
  python bench.py --workload config3 --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2a_ncu4.log 2>&1
ls -la $O | tail -20
