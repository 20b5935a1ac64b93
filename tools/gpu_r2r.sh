mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4
timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 2
echo "--- butterflies"
HEIC_B200_TRANSFORM_MMA=0 timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"transform|tu_list" -c 12 --csv --log-file gpurun_out/r2r_tr_launches.csv python tools/profile_batch.py --decodes 2 --stage-reps 0 > /dev/null 2>&1
HEIC_B200_TRANSFORM_MMA=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"transform|tu_list" -c 12 --csv --log-file gpurun_out/r2r_tr_launches_bf.csv python tools/profile_batch.py --decodes 2 --stage-reps 0 > /dev/null 2>&1
