timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 2
timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"transform|tu_list" -c 6 --csv --log-file gpurun_out/r2z_launches.csv python tools/profile_batch.py --decodes 1 --stage-reps 0 > /dev/null 2>&1
