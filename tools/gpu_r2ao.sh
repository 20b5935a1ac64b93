timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_synth.py -x -q -m gpu 2>&1 | tail -n 1
echo "--- default (deblock 12 CTAs)"
timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 1
for v in list5 list6 list8; do
echo "--- $v"
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_$v.so timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 1
done
