mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"transform|tu_list|deblock" -c 8 -o gpurun_out/r2s_tr_db python tools/profile_batch.py --decodes 1 --stage-reps 0 > gpurun_out/r2s_ncu.log 2>&1; echo "ncu rc=$?"
