timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 2
timeout 900 python bench.py --no-cpu --no-converged --steps 5 2> gpurun_out/r2ab.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ordered', d['value'], d['e2e'])"
HEIC_B200_PIPE_ORDER=0 timeout 900 python bench.py --no-cpu --no-converged --steps 5 2> gpurun_out/r2ab0.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unordered', d['value'], d['e2e'])"
