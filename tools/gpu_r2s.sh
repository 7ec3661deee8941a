#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:k_score_diff -c 1 -o $O/r2s_k_score_diff_config5 -f $B > $O/r2s_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -c 1 -o $O/r2s_k_scatter_config5 -f $B > $O/r2s_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -c 1 -o $O/r2s_k_scatter_config3 -f $B --workload config3 > $O/r2s_ncu3.log 2>&1
ls $O | grep r2s
