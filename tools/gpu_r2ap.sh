for v in intra72 intra96; do
echo "--- $v"
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_$v.so timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 1
done
