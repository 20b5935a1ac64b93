timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 2
timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:"cabac_kernel" -c 1 python tools/profile_batch.py --decodes 1 --stage-reps 0 2>&1 | grep -E "inst_executed|issue_active|duration"
