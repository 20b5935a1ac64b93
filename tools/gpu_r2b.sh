# round 2, second GPU call: parity + bench on the final pool (realistic SAO / split statistics), ncu of the CABAC kernel
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2b_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2b_bench.err
python tools/profile_batch.py --stage-reps 2 > gpurun_out/r2b_stages.log 2>&1; echo "stages rc=$?"; tail -n 3 gpurun_out/r2b_stages.log
python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches.csv python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2b_ncu1.log 2>&1; echo "ncu launches rc=$?"
python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2b_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cabac_kernel -s 1 -c 1 -o gpurun_out/r2b_cabac python tools/profile_batch.py --decodes 1 --stage-reps 1 > gpurun_out/r2b_ncu2.log 2>&1; echo "ncu cabac rc=$?"
