"""The opt-out paths selected by environment variables (read once per process, hence the subprocess): the register
butterflies for 8x8 .. 32x32 transforms (HEIC_B200_TRANSFORM_MMA=0, the A/B partner of the tensor-core kernels) and the
unordered chunk pipeline (HEIC_B200_PIPE_ORDER=0) must stay bit-exact too."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize(
    "env,target",
    [
        ({"HEIC_B200_TRANSFORM_MMA": "0"}, "tests/test_gpu_synth.py::test_stages_match_oracle"),
        ({"HEIC_B200_PIPE_ORDER": "0"}, "tests/test_gpu_pipeline.py"),
    ],
    ids=["transform_butterflies", "pipeline_unordered"],
)
def test_opt_out_path_is_bit_exact(built, env, target):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-m", "pytest", target, "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
