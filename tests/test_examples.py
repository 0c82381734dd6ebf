"""Every script under examples/ runs to completion on CPU (the reference smoke-tests its modelzoo the same way: cibuild/model-test.sh)."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPTS = sorted(s for s in glob.glob(os.path.join(ROOT, "examples", "*.py")) if not os.path.basename(s).startswith("_"))


@pytest.mark.parametrize("script", SCRIPTS, ids=[os.path.basename(s) for s in SCRIPTS])
def test_example_runs(script):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")          # no PYTHONPATH help: the scripts must run from a plain checkout
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
