mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2m_pytest.log
timeout 300 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -n 2
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_dbu.so timeout 300 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -n 2
