mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 2
timeout 1200 python bench.py --stages > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench rc=$?"
tail -c 400 gpurun_out/final_bench_default.json
