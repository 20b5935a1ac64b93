mkdir -p gpurun_out
for cap in 150 100 70; do
echo "--- cap $cap"
HEIC_B200_CABAC_GROUP_CAP=$cap timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | grep -E "groups|cabac" | cut -c1-200
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"intra_kernel" -s 1 -c 1 -o gpurun_out/r2u_intra python tools/profile_batch.py --decodes 1 --stage-reps 0 > gpurun_out/r2u_ncu.log 2>&1; echo "ncu rc=$?"
