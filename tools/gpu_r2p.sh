mkdir -p gpurun_out
timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 3 2>&1 | tail -n 3
echo "--- mbar"
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_mbar.so timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 3 2>&1 | tail -n 3
echo "--- mbar parity"
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_mbar.so timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 2
echo "--- trace"
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_trace.so timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 0 > gpurun_out/r2p_trace.log 2>&1; grep -c '^CT' gpurun_out/r2p_trace.log
