"""CPU checks: the C-ABI library builds, loads, and exports every symbol include/b2rl.h declares;
the product path fails loudly without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from agilerl_b200.csrc import build
    return build.build()


def test_header_symbols_exported(libpath):
    hdr = open(os.path.join(ROOT, "include", "b2rl.h")).read()
    declared = set(re.findall(r"\b(b2rl_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(libpath)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in b2rl.h but not exported: {missing}"
    from agilerl_b200 import _lib
    assert set(_lib.EXPORTS) <= declared | {"b2rl_last_error"}
    assert lib.b2rl_version() >= 100


def test_struct_layouts_match_header():
    from agilerl_b200 import _lib
    # sizes computed by the C compiler for the same structs
    import subprocess, tempfile, textwrap
    src = textwrap.dedent(f"""
        #include <stdio.h>
        #include "{ROOT}/include/b2rl.h"
        int main(void) {{ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b2rl_layer), sizeof(b2rl_net_desc),
                           sizeof(b2rl_learn_cfg), sizeof(b2rl_learn_bufs), sizeof(b2rl_step_state),
                           sizeof(b2rl_ddpg_cfg), sizeof(b2rl_ddpg_bufs), sizeof(b2rl_maddpg_cfg),
                           sizeof(b2rl_maddpg_bufs)); return 0; }}
    """)
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c"); exe = os.path.join(d, "s")
        open(c, "w").write(src)
        subprocess.check_call(["gcc", c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.Layer), ctypes.sizeof(_lib.NetDesc), ctypes.sizeof(_lib.LearnCfg),
                     ctypes.sizeof(_lib.LearnBufs), ctypes.sizeof(_lib.StepState), ctypes.sizeof(_lib.DdpgCfg),
                     ctypes.sizeof(_lib.DdpgBufs), ctypes.sizeof(_lib.MaddpgCfg), ctypes.sizeof(_lib.MaddpgBufs)]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from agilerl_b200 import _lib
    from agilerl_b200.components import PrioritizedReplayBuffer, ReplayBuffer
    for ctor in (lambda: ReplayBuffer(8, device="cpu"), lambda: PrioritizedReplayBuffer(8, device="cpu"),
                 lambda: ReplayBuffer(8, device="cuda")):
        with pytest.raises(_lib.B2RLError):
            ctor()
