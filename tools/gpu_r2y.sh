mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"cabac_kernel" -c 1 -o gpurun_out/r2y_cabac python tools/profile_batch.py --decodes 1 --stage-reps 0 > gpurun_out/r2y_ncu.log 2>&1; echo "ncu rc=$?"
