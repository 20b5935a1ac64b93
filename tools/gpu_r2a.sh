# round 2, first GPU call: parity on the default and the pair-by-position builds, bench on the pool workload, ncu of that batch
mkdir -p gpurun_out
POS=$PWD/heif_b200/variants/libheic_pos.so
python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest default rc=$?"; tail -2 gpurun_out/r2a_pytest.log
HEIC_B200_LIB=$POS python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_pos.log 2>&1; echo "pytest pos rc=$?"; tail -2 gpurun_out/r2a_pytest_pos.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
python tools/profile_batch.py --stage-reps 2 > gpurun_out/r2a_stages.log 2>&1; echo "stages rc=$?"; tail -3 gpurun_out/r2a_stages.log
HEIC_B200_LIB=$POS python tools/profile_batch.py --stage-reps 2 > gpurun_out/r2a_stages_pos.log 2>&1; echo "stages pos rc=$?"; tail -3 gpurun_out/r2a_stages_pos.log
HEIC_B200_CABAC_DEAL=1 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -3
python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2a_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches.csv python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2a_ncu1.log 2>&1; echo "ncu launches rc=$?"
python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2a_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cabac_kernel -s 1 -c 1 -o gpurun_out/r2a_cabac python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2a_ncu2.log 2>&1; echo "ncu cabac rc=$?"
