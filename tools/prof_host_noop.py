"""Host-side cost of the reference-shaped API with the C entry points replaced by no-ops — runs WITHOUT a GPU.

    python tools/prof_host_noop.py [--profile]

What it measures: interpreter + torch-CPU time of ``clone()`` / ``TournamentSelection.select`` / ``Mutations.mutation`` /
agent construction (pure host work: faithful), and of ingest / sample / update_priorities (indicative only: with the
"device" mapped to the CPU the staging path re-stages what would already be on the GPU).  It found the two host stalls
fixed at the end of round 2: the multi-threaded QR of the orthogonal initialisation (112 ms per conv layer here) and
``torch.randperm`` over the whole buffer in ``ReplayBuffer.sample``."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B2RL_GRAPH"] = "0"                      # graph capture needs the real library

import numpy as np  # noqa: E402
import torch  # noqa: E402

import agilerl_b200  # noqa: E402,F401
from agilerl_b200 import _lib  # noqa: E402
from agilerl_b200.components import replay_buffer as rb  # noqa: E402


class NoopLib:
    def __getattr__(self, name):
        if name == "b2rl_noise_count":
            def f(desc, out):
                out._obj.value = 64
                return 0
            return f
        if name.endswith("workspace_bytes"):
            def f(*a):
                a[-1]._obj.value = 256
                return 0
            return f
        if name == "b2rl_host_randperm_prefix":     # keep the real host helpers
            raise AttributeError(name)
        return lambda *a, **k: 0


class _Event:
    def record(self, *a): pass
    def synchronize(self): pass
    def wait(self, *a): pass
    def query(self): return True


_lib.as_device = lambda d: torch.device("cpu")
_lib.load = lambda require_cuda=False: NoopLib()
_lib.stream_ptr = lambda d=None: 0
_lib.check = lambda rc: None
_lib.require_cuda_tensor = lambda t, what="tensor": None
torch.Tensor.pin_memory = lambda self: self
torch.cuda.Event = _Event
rb._PinnedRing.sent = lambda self, k, dev: None

from agilerl_b200.algorithms import RainbowDQN  # noqa: E402
from agilerl_b200.compat import spaces  # noqa: E402
from agilerl_b200.components import MultiStepReplayBuffer, PrioritizedReplayBuffer, Transition  # noqa: E402
from agilerl_b200.hpo import Mutations, TournamentSelection  # noqa: E402

NET = {"encoder_config": {"channel_size": [32, 32], "kernel_size": [8, 4], "stride_size": [4, 2]},
       "head_config": {"hidden_size": [64]}, "latent_dim": 32}


def us(fn, n):
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t) / n * 1e6


def main():
    mk = lambda i: RainbowDQN(spaces.Box(0, 255, (4, 84, 84), np.uint8), spaces.Discrete(6), index=i, net_config=dict(NET),
                              batch_size=256, v_min=-10.0, v_max=10.0)
    mk(98); mk(99)                                   # the first two constructions pay one-off library warm-ups (~1 s)
    print(f"construct one agent      {us(lambda: mk(9), 3) / 1e3:8.2f} ms")
    pop = [mk(i) for i in range(8)]
    print(f"clone()                  {us(lambda: pop[0].clone(), 5) / 1e3:8.2f} ms")
    for a in pop:
        a.fitness = [float(np.random.rand())]
    t = time.perf_counter(); _, new = TournamentSelection(2, True, 8, 1).select(pop)
    print(f"select (8 members)       {(time.perf_counter() - t) * 1e3:8.2f} ms")
    t = time.perf_counter(); new = Mutations(0.2, 0.2, 0.2, 0.2, 0.1, 0.1, rand_seed=0, device="cuda").mutation(new)
    print(f"mutation (8 members)     {(time.perf_counter() - t) * 1e3:8.2f} ms   {[a.mut for a in new]}")

    E, g = 4, np.random.default_rng(0)
    host = dict(obs=g.integers(0, 256, (E, 4, 84, 84), dtype=np.uint8), action=g.integers(0, 6, (E,)),
                reward=g.standard_normal(E).astype(np.float32), next_obs=g.integers(0, 256, (E, 4, 84, 84), dtype=np.uint8),
                done=np.zeros(E, bool))
    mem, nmem = PrioritizedReplayBuffer(4096, 0.6, device="cuda"), MultiStepReplayBuffer(4096, 3, 0.99, device="cuda")

    def ingest():
        one = nmem.add(Transition(**host, batch_size=[E]).to_tensordict())
        if one is not None:
            mem.add(one)

    def sample():
        exp = mem.sample(256, 0.4)
        nmem.sample_from_indices(exp["idxs"].squeeze(1))
        return exp
    for _ in range(50):
        ingest()
    exp, pri = sample(), np.abs(g.standard_normal(256)).astype(np.float32)
    print(f"ingest (indicative)      {us(ingest, 1000):8.1f} us")
    print(f"sample (indicative)      {us(sample, 300):8.1f} us")
    print(f"update_priorities        {us(lambda: mem.update_priorities(exp['idxs'], pri), 1000):8.1f} us")
    if "--profile" in sys.argv:
        import cProfile
        import pstats
        pr = cProfile.Profile(); pr.enable()
        for _ in range(300):
            ingest(); mem.update_priorities(sample()["idxs"], pri)
        pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(30)


if __name__ == "__main__":
    main()
