# Reduced final pass after the last kernel changes: default bench line, launch list of the bench command, --set full of one decode
mkdir -p gpurun_out
timeout 1200 python bench.py --stages > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench rc=$?"
tail -c 300 gpurun_out/final_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-converged > gpurun_out/final_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -c 12 -o gpurun_out/final_full python tools/profile_batch.py --decodes 1 --stage-reps 0 > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
