# The round's final measurement pass: parity suite, both bench arms, the ncu launch list of the bench command and one
# `--set full` capture of every kernel of a decode.  Everything lands in gpurun_out/final_*.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/final_pytest.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?"
timeout 1200 python bench.py --stages > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench rc=$?"
tail -c 600 gpurun_out/final_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-converged > gpurun_out/final_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -c 16 -o gpurun_out/final_full python tools/profile_batch.py --decodes 1 --stage-reps 0 > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
