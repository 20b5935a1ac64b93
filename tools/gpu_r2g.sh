mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2g_pytest.log
timeout 500 python bench.py --no-cpu --steps 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2g_bench.err
timeout 300 python tools/profile_batch.py --batch 96 --decodes 1 --stage-reps 1 > gpurun_out/r2g_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:intra_kernel -s 1 -c 1 -o gpurun_out/r2g_intra python tools/profile_batch.py --batch 96 --decodes 1 --stage-reps 1 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu intra rc=$?"
