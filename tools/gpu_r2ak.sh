for pct in 20 35 50; do
echo "--- filler $pct"
HEIC_B200_CABAC_FILLER=$pct timeout 300 python tools/profile_batch.py --decodes 1 --stage-reps 2 2>&1 | tail -n 2
done
