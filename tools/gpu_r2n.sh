mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2n_pytest.log
timeout 300 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -n 2
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_c4.so timeout 300 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -n 2
