# 8 GPUs of one box: raw concurrent D2H probe, then the scaling bench at N = 8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2k_topo.txt 2>&1; free -g | head -2; nproc
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/d2h_probe.py > gpurun_out/r2k_d2h_probe_8.json 2> gpurun_out/r2k_probe.err; echo "probe rc=$?"; cat gpurun_out/r2k_d2h_probe_8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2k_bench_8gpu.json 2> gpurun_out/r2k_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/r2k_bench_8gpu.err; tail -c 1500 gpurun_out/r2k_bench_8gpu.json
