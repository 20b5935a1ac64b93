# gpurun_out/final_* (tools/gpu_final.sh) -> the tracked artefacts under profiles/r02_final_*
set -e
cd "$(dirname "$0")/.."
O=gpurun_out P=profiles
tail -n 1 $O/final_bench_default.json > $P/r02_bench_default_run.json
tail -n 1 $O/final_bench_reference.json > $P/r02_bench_reference_run.json
grep -v '^==' $O/final_launches.csv > $P/r02_final_launches.csv
python tools/ncu_summary.py $O/final_full.ncu-rep > $P/r02_final_ncu_full_summary.txt
python tools/profile_digest.py $P/r02_final_launches.csv $P/r02_final_ncu_full_summary.txt 592 60 r02_final > /dev/null
for k in cabac_kernel intra_kernel transform_mma_kernel deblock_kernel; do
  ncu -i $O/final_full.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:$k > /tmp/_src.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/_src.csv 40 > $P/r02_final_${k}_hotspots.txt
done
{
  echo "# cuobjdump -sass heif_b200/libheic_b200.so: instruction classes that show the Blackwell-specific paths (count of SASS lines)"
  for pat in 'IMMA\.16832' 'IMMA\.16816' 'LDSM' 'UBLKCP' 'VIADDMNMX' 'LDGSTS' 'I2IP' 'REDUX'; do
    printf "%-14s %s\n" "$pat" "$(cuobjdump -sass heif_b200/libheic_b200.so 2>/dev/null | grep -cE "$pat")"
  done
  echo; echo "# examples"
  cuobjdump -sass heif_b200/libheic_b200.so 2>/dev/null | grep -E 'IMMA|LDSM|UBLKCP|VIADDMNMX|I2IP' | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | awk '{$1=""; print}' | sort | uniq -c | sort -rn | head -24
} > $P/r02_sass_evidence.txt
echo done
