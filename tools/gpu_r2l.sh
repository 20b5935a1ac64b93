mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2l_pytest.log
timeout 300 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -n 2
timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2l_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2l_launches.csv python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2l_ncu1.log 2>&1; echo "ncu launches rc=$?"
