mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2f_pytest.log
timeout 400 python bench.py --no-cpu --no-converged --steps 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2f_bench.err
HEIC_B200_TRACE=2 timeout 300 python bench.py --no-cpu --no-converged --batch 256 --e2e-batch 256 --steps 3 --check-images 2 > gpurun_out/r2f_trace.json 2> gpurun_out/r2f_trace.err; echo "trace rc=$?"; grep -c chunk gpurun_out/r2f_trace.err
