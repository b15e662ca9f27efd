"""pytest plugin used by tests/test_reference_suite_cpu.py: makes the REFERENCE's OWN unit tests (run in place from
/root/reference/tests, unmodified) exercise THIS package's classes.  On import: oracle.refshim (the reference importable,
stand-ins for its missing third-party dependencies), Python stand-ins for the C entry points those tests reach (no GPU
here: bytes movers, the n-step fold, and the priority trees through the oracle's C segment tree operating on the buffers
our classes own), then ``agilerl_b200.install()`` so ``from agilerl.components... import ...`` binds our classes."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("B2RL_GRAPH", "0")

from oracle import refshim  # noqa: E402

refshim.install()
import agilerl_b200  # noqa: E402
from agilerl_b200 import _lib  # noqa: E402
from agilerl_b200.components import replay_buffer as rb  # noqa: E402
from oracle.segtree import load_lib as _oracle_tree  # noqa: E402
from test_multi_agent_host_cpu import StandIn, _bytes, _f32, _i64  # noqa: E402

OT = _oracle_tree()


def _f64(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_double * n).from_address(ptr))


class Lib(StandIn):
    def b2rl_ring_write(self, dst, src, row_bytes, start, n, max_size, stream):
        return self.b2rl_ring_write_multi(1, [dst], [src], [row_bytes], start, n, max_size, stream)

    def b2rl_gather_rows(self, dst, src, idx, row_bytes, n, stream):
        return self.b2rl_gather_rows_multi(1, [dst], [src], [row_bytes], idx, n, stream)

    @staticmethod
    def _stop(done_steps, n, E):
        last = 0
        for k in range(1, n):
            last = k
            if np.any(_f32(done_steps[k], E) != 0):
                break
        return last

    def b2rl_nstep_fold(self, reward_steps, done_steps, n, E, gamma, reward_out, last_out, stream):
        last = self._stop(done_steps, n, E)
        r = _f32(reward_steps[0], E).copy()
        for k in range(1, last + 1):
            r += _f32(reward_steps[k], E) * np.float32(gamma ** k)
        _f32(reward_out, E)[:] = r
        np.ctypeslib.as_array((ctypes.c_int32 * 1).from_address(last_out))[0] = last
        return 0

    def b2rl_select_copy(self, dst, srcs, n, which, nbytes, stream):
        k = int(np.ctypeslib.as_array((ctypes.c_int32 * 1).from_address(which))[0])
        _bytes(dst, nbytes)[:] = _bytes(srcs[k], nbytes)
        return 0

    def b2rl_nstep_ingest(self, nf, ring, src, row_bytes, role, rew, don, n, E, gamma, cursor, max_size, stream):
        last = self._stop(don, n, E)
        for f in range(nf):
            rbs = row_bytes[f]
            d = _bytes(ring[f], rbs * max_size)
            if role[f] == 2:
                r = _f32(rew[0], E).copy()
                for k in range(1, last + 1):
                    r += _f32(rew[k], E) * np.float32(gamma ** k)
                rows = r.view(np.uint8).reshape(E, 4)
            else:
                rows = _bytes(src[f * n + (last if role[f] == 1 else 0)], rbs * E).reshape(E, rbs)
            for e in range(E):
                slot = (cursor + e) % max_size
                d[slot * rbs:(slot + 1) * rbs] = rows[e]
        return 0

    def b2rl_tree_init(self, s, m, cap, stream):
        OT.ost_init(s, cap, 0); OT.ost_init(m, cap, 1)
        return 0

    def b2rl_tree_set(self, s, m, cap, idx, pa, n, stream):
        for i, x in zip(_i64(idx, n), _f64(pa, n)):
            if s:
                OT.ost_set(s, cap, 0, int(i), float(x))
            if m:
                OT.ost_set(m, cap, 1, int(i), float(x))
        return 0

    def b2rl_tree_set_range(self, s, m, cap, tree_ptr, n, max_size, p_alpha, stream):
        for k in range(n):
            i = (tree_ptr + k) % max_size
            OT.ost_set(s, cap, 0, i, p_alpha); OT.ost_set(m, cap, 1, i, p_alpha)
        return 0

    def b2rl_tree_retrieve(self, s, cap, ub, n, out, stream):
        u, o = _f64(ub, n), _i64(out, n)
        for k in range(n):
            o[k] = OT.ost_retrieve(s, cap, float(u[k]))
        return 0

    def b2rl_per_sample(self, s, m, cap, u, B, beta, size, idx_out, w_out, stream):
        OT.oper_sample(s, cap, u, B, idx_out)
        if w_out:
            OT.oper_weights(s, m, cap, idx_out, B, beta, size, w_out)
        return 0

    def b2rl_host_priority_pow(self, pri, n, alpha, floor_, out, mx):
        p = np.maximum(_f32(pri, n).astype(np.float64), floor_)
        _f64(out, n)[:] = [float(x) ** alpha for x in p]
        mx._obj.value = max(mx._obj.value, float(p.max()))
        return 0


    # construction of the value-based agents only asks for the size of their noise vector
    def b2rl_noise_count(self, desc, out):
        d = desc._obj
        out._obj.value = sum(layers[i].in_c + layers[i].out_c for layers, cnt in ((d.enc, d.n_enc), (d.val, d.n_val), (d.adv, d.n_adv))
                             for i in range(cnt) if layers[i].noisy)
        return 0


    def _noop(self, *a):
        return 0
    b2rl_noise_reset_philox = b2rl_noise_reset_from_normals = _noop          # fresh noise at construction / clone


class _Event:
    def record(self, *a): pass
    def synchronize(self): pass
    def wait(self, *a): pass
    def query(self): return True


_lib.as_device = lambda d: torch.device("cpu")
_lib.load = lambda require_cuda=False, _l=Lib(): _l
_lib.stream_ptr = lambda d=None: 0
_lib.check = lambda rc: None
_lib.require_cuda_tensor = lambda t, what="tensor": None
torch.Tensor.pin_memory = lambda self: self
torch.cuda.Event = _Event
rb._PinnedRing.sent = lambda self, k, dev: None
agilerl_b200.install(include_driver=False)

# tests/test_hpo/test_tournament.py imports a helper of the LLM tests at module level (transformers / peft / accelerate): an
# inert stand-in keeps the file importable; the LLM cases themselves are outside this package and fail as expected
import types  # noqa: E402

_grpo = types.ModuleType("tests.test_algorithms.test_llms.test_grpo")
_grpo.create_module = lambda *a, **k: None
sys.modules.setdefault("tests.test_algorithms.test_llms.test_grpo", _grpo)
