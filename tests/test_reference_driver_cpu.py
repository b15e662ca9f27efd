"""CPU, build container only: the UNCHANGED reference driver (/root/reference/agilerl/training/train_off_policy.py, run
through oracle.refshim) trains a population of THIS package's agents bound under the ``agilerl.*`` names by
``agilerl_b200.install()`` — BASELINE configs[0] (DQN, pop = 4, vector environment) with tournament selection and mutation.
The C entry points are Python stand-ins (tests/_reference_driver_standin.py): the run pins the drop-in claim at the call
level; the kernels behind those entry points are the GPU tests' business."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/agilerl"), reason="needs the reference source tree")


def _run(script):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-3000:])
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
    assert line, out.stdout[-1500:]
    return json.loads(line[-1][len("RESULT "):])


def test_unchanged_reference_driver_trains_our_dqn_population():
    r = _run("_reference_driver_standin.py")
    assert r["pop"] == 4 and r["types"] == ["agilerl_b200.algorithms.dqn"]
    assert r["generations"] == 3 and r["fit_width"] == [4, 4, 4]              # 120 steps / 40 per generation
    assert all(s >= 120 for s in r["steps"]) and all(n >= 1 for n in r["fitness_len"])
    # 4 agents x 3 generations x 20 vector steps, every one stored; learning every LEARN_STEP / num_envs steps once the
    # buffer holds a batch
    assert r["calls"]["adds"] == 240 and r["memory_len"] == 480
    assert 200 <= r["calls"]["learn"] <= 240 and r["calls"]["forward_rows"] > 480
