mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-converged > gpurun_out/final_bench_8gpu.json 2> gpurun_out/final_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/final_bench_8gpu.err; tail -c 1200 gpurun_out/final_bench_8gpu.json
