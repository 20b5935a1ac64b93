mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2o_pytest.log
timeout 300 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -n 2
timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2o_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"cabac_kernel|intra_kernel" -s 2 -c 2 -o gpurun_out/r2o_cabac_intra python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2o_ncu.log 2>&1; echo "ncu rc=$?"
