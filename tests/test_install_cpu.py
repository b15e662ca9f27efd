"""CPU: ``agilerl_b200.install()`` puts this package's modules under the reference's names, so that the UNCHANGED
``agilerl/training/train_off_policy.py`` (executed here from /root/reference through oracle.refshim; skipped on the
GPU box, where the reference does not exist) binds OUR classes: its ``Sampler``'s ``isinstance`` dispatch
(sampler.py:71-72) then recognises our buffers."""
import importlib
import inspect
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_install_without_reference_creates_namespace_and_maps_driver():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import agilerl_b200\n"
        "names = agilerl_b200.install()\n"
        "import agilerl.components.replay_buffer as rb, agilerl.algorithms as alg, agilerl.training.train_off_policy as t\n"
        "import agilerl_b200.components.replay_buffer as ours\n"
        "assert rb is ours and alg.RainbowDQN.__module__.startswith('agilerl_b200')\n"
        "from agilerl.components import PrioritizedReplayBuffer\n"
        "from agilerl.components.sampler import Sampler\n"
        "assert PrioritizedReplayBuffer is ours.PrioritizedReplayBuffer\n"
        "assert t.__name__ == 'agilerl_b200.training.train_off_policy' and 'agilerl.hpo.tournament' in names\n"
        "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/agilerl"), reason="needs the reference source tree")
def test_unchanged_reference_driver_binds_our_classes():
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import refshim; refshim.install()\n"
        "import agilerl_b200\n"
        "agilerl_b200.install(include_driver=False)\n"
        "import agilerl.training.train_off_policy as T, inspect\n"
        "assert inspect.getsourcefile(T).startswith('/root/reference/'), inspect.getsourcefile(T)\n"
        "import agilerl_b200.components as C, agilerl_b200.algorithms as A, agilerl_b200.hpo as H\n"
        "assert T.PrioritizedReplayBuffer is C.PrioritizedReplayBuffer and T.MultiStepReplayBuffer is C.MultiStepReplayBuffer\n"
        "assert T.Sampler is C.Sampler and T.Transition is C.Transition\n"
        "assert T.RainbowDQN is A.RainbowDQN and T.DQN is A.DQN and T.TD3 is A.TD3 and T.DDPG is A.DDPG\n"
        "assert T.TournamentSelection is H.TournamentSelection and T.Mutations is H.Mutations\n"
        "import inspect as i\n"
        "sig = i.signature(T.train_off_policy)\n"
        "assert list(sig.parameters)[:6] == ['env', 'env_name', 'algo', 'pop', 'memory', 'INIT_HP']\n"
        "print('ok')\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, (out.stdout[-500:], out.stderr[-3000:])
