mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2j_pytest.log
HEIC_B200_TRACE=2 timeout 600 python bench.py --no-cpu --steps 3 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; grep -v "chunk\|submit" gpurun_out/r2j_bench.err | tail -c 300
