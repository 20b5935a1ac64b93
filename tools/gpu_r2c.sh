# FSM CABAC: parity suite, then A/B timing against the nested walker on the pool batch
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2c_pytest.log
timeout 300 python tools/profile_batch.py --stage-reps 2 > gpurun_out/r2c_stages_fsm.log 2>&1; echo "fsm rc=$?"; tail -n 3 gpurun_out/r2c_stages_fsm.log
HEIC_B200_CABAC_FSM=0 timeout 300 python tools/profile_batch.py --stage-reps 2 > gpurun_out/r2c_stages_nested.log 2>&1; echo "nested rc=$?"; tail -n 3 gpurun_out/r2c_stages_nested.log
