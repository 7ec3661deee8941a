#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-extra --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2h_launches_config5.csv $B > $O/r2h_ncu0.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r2h_launches_config5_fp32.csv $B --precision fp32_fast > $O/r2h_ncu0b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -c 1 -o $O/r2h_k_scatter_config5 -f $B > $O/r2h_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -c 1 -o $O/r2h_k_scatter_config5_fp32_u32 -f $B --precision fp32_fast > $O/r2h_ncu2.log 2>&1
SCDC_U=16 ncu --set full --clock-control none --import-source on -k regex:k_scatter -c 1 -o $O/r2h_k_scatter_config5_fp32_u16 -f $B --precision fp32_fast > $O/r2h_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pass1_cond -c 1 -o $O/r2h_k_pass1_config5 -f $B > $O/r2h_ncu4.log 2>&1
ls $O | grep r2h
