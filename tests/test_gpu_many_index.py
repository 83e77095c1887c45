"""Both index builders of the second-generation N-way union (rb200_many2.cu) against the reference:
the library picks key windows for long directories and per-container atomics for short ones, so the
or_many parity tests are re-run in a child process with each choice FORCED (RB200_OR_MANY_INDEX is
read once per process)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("index", ["window", "atomic"])
def test_or_many_parity_with_forced_index(index):
    env = dict(os.environ, RB200_OR_MANY_INDEX=index)
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "-k", "or_many",
           "tests/test_gpu_parity.py", "tests/test_gpu_sharded.py", "tests/test_gpu_properties.py"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", f"many_index_{index}.log"), "w") as f:
            f.write(r.stdout[-20000:] + r.stderr[-5000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
