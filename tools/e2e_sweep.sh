# e2e sweep over pipeline knobs: bash tools/e2e_sweep.sh  (prints e2e MP/s, ms per call, synchronous-call MP/s)
run() {
  echo "$*: $(env "$@" timeout 300 python bench.py --no-cpu --batch 256 --e2e-batch 256 --steps 6 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["e2e"]["value"], d["e2e"]["ms_per_call"], d["e2e"]["synchronous_call_MPps"])')"
}
run HEIC_B200_PIPE_CHUNK=256 HEIC_B200_PIPE_SLOTS=2
run HEIC_B200_PIPE_CHUNK=256 HEIC_B200_PIPE_SLOTS=3
run HEIC_B200_PIPE_CHUNK=128 HEIC_B200_PIPE_SLOTS=3
run HEIC_B200_PIPE_CHUNK=128 HEIC_B200_PIPE_SLOTS=6
run HEIC_B200_TRACE=1 HEIC_B200_PIPE_CHUNK=256 HEIC_B200_PIPE_SLOTS=3
