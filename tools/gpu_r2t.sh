mkdir -p gpurun_out
echo "--- ctas4"
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_ctas4.so timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 3 2>&1 | tail -n 3
