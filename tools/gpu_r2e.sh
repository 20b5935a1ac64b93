# nested CABAC with 4 CTAs/SM; e2e pipeline sweep (chunk x slots)
mkdir -p gpurun_out
HEIC_B200_LIB=$PWD/heif_b200/variants/libheic_ctas4.so timeout 200 python tools/profile_batch.py --stage-reps 2 2>&1 | tail -n 2
run() {
  echo "$*: $(env "$@" timeout 300 python bench.py --no-cpu --no-converged --batch 256 --e2e-batch 256 --steps 6 --check-images 2 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["e2e"]["value"], d["e2e"]["ms_per_call"], d["e2e"]["synchronous_call_MPps"], d["value"])')"
}
run HEIC_B200_PIPE_CHUNK=32 HEIC_B200_PIPE_SLOTS=8
run HEIC_B200_PIPE_CHUNK=32 HEIC_B200_PIPE_SLOTS=16
run HEIC_B200_PIPE_CHUNK=16 HEIC_B200_PIPE_SLOTS=16
run HEIC_B200_PIPE_CHUNK=64 HEIC_B200_PIPE_SLOTS=8
run HEIC_B200_PIPE_CHUNK=64 HEIC_B200_PIPE_SLOTS=4
run HEIC_B200_TRACE=1 HEIC_B200_PIPE_CHUNK=32 HEIC_B200_PIPE_SLOTS=12
