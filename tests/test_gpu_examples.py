"""The example scripts (non-interactive counterparts of the reference's Example/*.py) run to completion on the GPU."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("script", ["three_link_tracking.py", "robotic_arm_log_barrier.py", "vehicle_and_cart_pole.py",
                                    "result_stream.py"])
def test_example_runs(script, tmp_path):
    env = dict(os.environ, EXAMPLE_BATCH="512")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script)], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "nan" not in out.stdout.lower()
