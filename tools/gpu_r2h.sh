mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2h_pytest.log
timeout 300 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -n 2
HEIC_B200_TRACE=2 timeout 500 python bench.py --no-cpu --no-converged --steps 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; grep -v chunk gpurun_out/r2h_bench.err | tail -c 300
