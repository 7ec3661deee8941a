#!/bin/bash
# final round-2 validation on one GPU: GPU tests, smoke, reference arm, default bench, launch list, one-shot timing, ncu of pass 1 / labels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
( time python -m pytest tests -m gpu -q --durations=5 ) > $O/r2w_pytest.txt 2>&1
echo "pytest rc=$?" >> $O/r2w_pytest.txt
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/r2w_smoke.txt 2>&1
( time python bench.py --impl reference ) > $O/r2w_ref.json 2> $O/r2w_ref.err
( time python bench.py ) > $O/r2w_bench.json 2> $O/r2w_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2w_launches_config5.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --no-e2e > $O/r2w_ncu0.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pass1_cond -c 1 -o $O/r2w_k_pass1_cond_config5 -f $B > $O/r2w_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_draw_labels_keys2 -c 1 -o $O/r2w_k_draw_labels_keys2_config5 -f $B > $O/r2w_ncu2.log 2>&1
tail -3 $O/r2w_pytest.txt; tail -2 $O/r2w_smoke.txt
