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
  for pat in 'IMMA\.16832' 'IMMA\.16816' 'LDSM' 'IDP\.4A' 'UBLKCP' 'VIADDMNMX' 'LDGSTS' 'I2IP' 'REDUX'; do
    printf "%-14s %s\n" "$pat" "$(cuobjdump -sass heif_b200/libheic_b200.so 2>/dev/null | grep -cE "$pat")"
  done
  echo; echo "# examples"
  cuobjdump -sass heif_b200/libheic_b200.so 2>/dev/null | grep -E 'IMMA|LDSM|IDP|UBLKCP|VIADDMNMX|I2IP' | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | awk '{$1=""; print}' | sort | uniq -c | sort -rn | head -24
} > $P/r02_sass_evidence.txt

# ncu returns no DRAM counters for two kernels in the whole-decode capture (tu_list_kernel, deblock_kernel<0>): their bytes
# come from the dedicated captures committed as profiles/r02e_*
python - <<'PY'
import json
p = "profiles/dram_traffic_per_image.json"
d = json.load(open(p))
if "_supplemented" not in d and any("tu_list" in k for k in d.get("_not_captured", [])):
    d["transform"] += (1.893846e9 + 1.457417e9) / 592
if "_supplemented" not in d and any("deblock_kernel<0>" in k for k in d.get("_not_captured", [])):
    d["deblock"] += (9.435776e9 + 5.526232e9) / 592
d["_supplemented"] = ("ncu returned no DRAM counters for tu_list_kernel and deblock_kernel<0> in the whole-decode capture; their bytes are "
                      "taken from the dedicated captures of the same kernels (profiles/r02e_transform_deblock_before_ncu.txt: tu_list 1.89 + "
                      "1.46 GB; profiles/r02e_deblock_dp4a_ncu.txt: deblock<0> 9.44 + 5.53 GB per 592 images)")
json.dump(d, open(p, "w"), indent=1)
PY
echo done
