mkdir -p gpurun_out
timeout 300 python tools/profile_batch.py --stage-reps 2 > gpurun_out/r2d_stages_fsm.log 2>&1; echo "fsm rc=$?"; tail -n 3 gpurun_out/r2d_stages_fsm.log
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2d_pytest.log
timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2d_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cabac_fsm_kernel -s 1 -c 1 -o gpurun_out/r2d_cabac_fsm python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2d_ncu2.log 2>&1; echo "ncu cabac rc=$?"
