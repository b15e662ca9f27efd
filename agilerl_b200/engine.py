"""Device-side state and launch logic of one learning agent (host glue over libb2rl.so).

``NetBuffers`` = one network's flat fp32 parameter + epsilon buffers with named views that carry
the reference's ``state_dict()`` keys.  ``LearnEngine`` = actor + target + Adam moments + scratch,
and the calls that replace ``RainbowDQN.learn`` / ``DQN.learn`` bodies
(agilerl/algorithms/dqn_rainbow.py:369-490, dqn.py:326-347).
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict

import torch

from . import _lib
from .networks.spec import FlatLayout


class NetBuffers:
    def __init__(self, layout: FlatLayout, device):
        self.layout = layout
        self.device = _lib.as_device(device)
        self.params = torch.zeros(max(layout.n_params, 1), dtype=torch.float32, device=self.device)
        self.eps = torch.zeros(max(layout.n_eps, 1), dtype=torch.float32, device=self.device)

    def view(self, key: str) -> torch.Tensor:
        e = self.layout.entries[key]
        n = 1
        for s in e.shape:
            n *= s
        buf = self.params if e.buf == "param" else self.eps
        return buf[e.offset:e.offset + n].view(e.shape)

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return OrderedDict((k, self.view(k).clone()) for k in self.layout.entries)

    def load_state_dict(self, sd, strict: bool = True) -> None:
        for k in self.layout.entries:
            if k in sd:
                v = sd[k]
                v = v if isinstance(v, torch.Tensor) else torch.as_tensor(v)
                self.view(k).copy_(v.to(self.device, dtype=torch.float32))
            elif strict:
                raise KeyError(f"missing key {k} in state_dict")

    def copy_from(self, other: "NetBuffers") -> None:
        self.params.copy_(other.params)
        self.eps.copy_(other.eps)


import os as _os

_SIDE = int(_os.environ["B2RL_SIDE_STREAMS"]) if "B2RL_SIDE_STREAMS" in _os.environ else None
_GRAPH = _os.environ.get("B2RL_GRAPH", "1") != "0"      # CUDA-graph replay of the fused step (0: always eager)
_PREP = _os.environ.get("B2RL_PREP", "1") != "0"        # parameter-only work of a captured step on a side stream under the sampler


class _FusedPlan:
    """Static buffers and the two captured graphs of one fused-step configuration of an engine."""

    def __init__(self, eng, per, nmem, B, support, hp, gamma_n, weights_mode, side_streams):
        dev = eng.device
        self.eng, self.per, self.nmem, self.B = eng, per, nmem, B
        self.support, self.hp, self.gamma_n = support, dict(hp), gamma_n
        self.weights_mode, self.side_streams = weights_mode, side_streams
        self.idx = torch.empty(B, dtype=torch.int64, device=dev)
        self.w, self.a, self.r, self.d = (torch.empty(B, dtype=torch.float32, device=dev) for _ in range(4))
        self.loss_elem = torch.empty(B, dtype=torch.float32, device=dev)
        self.out = torch.empty(B + 1, dtype=torch.float32, device=dev)          # priorities | scalar loss
        self.state_host = _lib.StepState()
        self.state_dev = torch.zeros(ctypes.sizeof(_lib.StepState), dtype=torch.uint8, device=dev)
        self.front = self.tail = None
        self.fwd_done, self.done = torch.cuda.Event(), torch.cuda.Event()
        self.kernels = 0

    def capture(self) -> None:
        eng, per, nmem, B = self.eng, self.per, self.nmem, self.B
        lib = eng.lib
        desc = ctypes.byref(eng.layout.desc)
        f = nmem._fields
        dk = nmem.done_key or "done"
        batch = dict(obs=f[("obs",)], next_obs=f[(nmem.ns_key,)], action=self.a, reward=self.r, done=self.d)
        hp = self.hp
        cfg = eng._cfg(B, gamma=self.gamma_n, v_min=hp["v_min"], v_max=hp["v_max"], delta_z=hp["delta_z"],
                       weights_mode=self.weights_mode, driver_shapes=0, clip=1, lr=hp["lr"], tau=hp["tau"],
                       prior_eps=hp["prior_eps"], accumulate=0, use_noise=1, step=1, side_streams=self.side_streams)
        bufs, keep = eng._bufs(B, batch, self.w, self.support, self.loss_elem, self.out[:B], self.out[B:], None, self.idx)
        bufs.step_state = self.state_dev.data_ptr()
        self._keep = (keep, cfg, bufs)
        cap = torch.cuda.Stream(device=eng.device)
        cap.wait_stream(torch.cuda.current_stream(eng.device))
        s = cap.cuda_stream
        gh = ctypes.c_void_p()
        # ---- front: state write -> sample -> forwards/projection/loss -> priority write-back
        _lib.check(lib.b2rl_graph_begin(s))
        try:
            _lib.check(lib.b2rl_step_state_write(ctypes.byref(self.state_host), self.state_dev.data_ptr(), s))
            if _PREP:      # parameter-only work of the loss pass on a side stream, under the sampler
                _lib.check(lib.b2rl_rainbow_prep(desc, ctypes.byref(cfg), ctypes.byref(bufs), s))
                cfg.reserved_ = 1
            _lib.check(lib.b2rl_per_sample_fused_state(
                per.sum_tree.data_ptr, per.min_tree.data_ptr, per._cap, per._philox_seed, self.state_dev.data_ptr(), B,
                f[("action",)].data_ptr(), f[(nmem.reward_key,)].data_ptr(), f[(dk,)].data_ptr(), self.idx.data_ptr(),
                self.w.data_ptr(), self.a.data_ptr(), self.r.data_ptr(), self.d.data_ptr(), s))
            _lib.check(lib.b2rl_rainbow_loss(desc, ctypes.byref(cfg), ctypes.byref(bufs), s))
            _lib.check(lib.b2rl_tree_set_from_priorities(
                per.sum_tree.data_ptr, per.min_tree.data_ptr, per._cap, self.idx.data_ptr(), self.out.data_ptr(), B,
                float(per.alpha), 1e-5, per._max_priority_dev.data_ptr(), s))
        finally:
            _lib.check(lib.b2rl_graph_end(s, ctypes.byref(gh)))
        self.front = gh.value
        # ---- tail: backward -> clip/Adam/Polyak -> noise reset (actor, then target: dqn_rainbow.py:484-485)
        gt = ctypes.c_void_p()
        _lib.check(lib.b2rl_graph_begin(s))
        try:
            _lib.check(lib.b2rl_rainbow_backward(desc, ctypes.byref(cfg), ctypes.byref(bufs), s))
            _lib.check(lib.b2rl_optim_step(desc, ctypes.byref(cfg), ctypes.byref(bufs), s))
            _lib.check(lib.b2rl_noise_reset_state_pair(desc, eng.actor.eps.data_ptr(), eng.target.eps.data_ptr(),
                                                       eng.philox_seed, self.state_dev.data_ptr(), s))
        finally:
            _lib.check(lib.b2rl_graph_end(s, ctypes.byref(gt)))
        self.tail = gt.value
        n = ctypes.c_int(0)
        for g in (self.front, self.tail):
            _lib.check(lib.b2rl_graph_kernel_count(g, ctypes.byref(n)))
            self.kernels += n.value

    def destroy(self) -> None:
        lib = self.eng.lib
        for g in (self.front, self.tail):
            if g:
                lib.b2rl_graph_destroy(g)
        self.front = self.tail = None


class _ApiPlan:
    """Captured graphs of ``rainbow_learn`` on ONE set of batch buffers (keyed by their device addresses).

    ``agent.learn(experiences, ...)`` receives fresh tensors every step, but in a steady training loop the
    caching allocator hands the sampler the same blocks again and again: when every pointer of a call matches a
    plan, the step is two graph launches (forwards + projection + loss + D2H of loss/priorities; backward + optimiser
    + noise reset) instead of ~45 eager launches; any other call runs eagerly.  Nothing is copied, nothing is
    assumed about the tensors beyond their addresses."""

    def __init__(self, eng, B):
        dev = eng.device
        self.eng, self.B = eng, B
        self.loss_elem = torch.empty(B, dtype=torch.float32, device=dev)
        self.out = torch.empty(B + 1, dtype=torch.float32, device=dev)
        self.host = torch.empty(B + 1, dtype=torch.float32).pin_memory()
        self.state_host = _lib.StepState()
        self.state_dev = torch.zeros(ctypes.sizeof(_lib.StepState), dtype=torch.uint8, device=dev)
        self.front = self.tail = None
        self.seen = 0
        self.rb_ev, self.fwd_done, self.done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

    def capture(self, batch, gamma, driver, weights, weights_mode, support, hp, side_streams) -> None:
        eng, B = self.eng, self.B
        lib = eng.lib
        desc = ctypes.byref(eng.layout.desc)
        cfg = eng._cfg(B, gamma=gamma, v_min=hp["v_min"], v_max=hp["v_max"], delta_z=hp["delta_z"], weights_mode=weights_mode,
                       driver_shapes=int(driver), clip=1, lr=hp["lr"], tau=hp["tau"], prior_eps=hp["prior_eps"], accumulate=0,
                       use_noise=1, step=1, side_streams=side_streams)
        bufs, keep = eng._bufs(B, batch, weights, support, self.loss_elem, self.out[:B], self.out[B:], None, None)
        bufs.step_state = self.state_dev.data_ptr()
        self._keep = (cfg, bufs)
        cap = torch.cuda.Stream(device=eng.device)
        cap.wait_stream(torch.cuda.current_stream(eng.device))
        s = cap.cuda_stream
        gh, gt = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(lib.b2rl_graph_begin(s))
        try:
            _lib.check(lib.b2rl_step_state_write(ctypes.byref(self.state_host), self.state_dev.data_ptr(), s))
            _lib.check(lib.b2rl_rainbow_loss(desc, ctypes.byref(cfg), ctypes.byref(bufs), s))
            _lib.check(lib.b2rl_copy_d2h(self.host.data_ptr(), self.out.data_ptr(), (B + 1) * 4, s))
        finally:
            _lib.check(lib.b2rl_graph_end(s, ctypes.byref(gh)))
        _lib.check(lib.b2rl_graph_begin(s))
        try:
            _lib.check(lib.b2rl_rainbow_backward(desc, ctypes.byref(cfg), ctypes.byref(bufs), s))
            _lib.check(lib.b2rl_optim_step(desc, ctypes.byref(cfg), ctypes.byref(bufs), s))
            _lib.check(lib.b2rl_noise_reset_state_pair(desc, eng.actor.eps.data_ptr(), eng.target.eps.data_ptr(),
                                                       eng.philox_seed, self.state_dev.data_ptr(), s))
        finally:
            _lib.check(lib.b2rl_graph_end(s, ctypes.byref(gt)))
        self.front, self.tail = gh.value, gt.value

    def destroy(self) -> None:
        lib = self.eng.lib
        for g in (self.front, self.tail):
            if g:
                lib.b2rl_graph_destroy(g)
        self.front = self.tail = None


class LearnEngine:
    def __init__(self, layout: FlatLayout, actor: NetBuffers, target: NetBuffers):
        self.layout, self.actor, self.target = layout, actor, target
        self.device = actor.device
        self.lib = _lib.load()
        n = max(layout.n_params, 1)
        self.grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.step = 0
        self._ws: dict[tuple, torch.Tensor] = {}
        self._noise_count = ctypes.c_int64(0)
        _lib.check(self.lib.b2rl_noise_count(ctypes.byref(layout.desc), ctypes.byref(self._noise_count)))
        self.philox_seed = 0xB200
        self._host_out: dict = {}
        self.philox_offset = 0
        self._plans: dict = {}
        self._api_plans: dict = {}

    def __del__(self):
        try:
            for plan in list(self._plans.values()) + list(self._api_plans.values()):
                plan.destroy()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # -- scratch -------------------------------------------------------------------------------
    # -- two-stream learn ----------------------------------------------------------------------
    # With ``overlap=True`` the backward + optimiser + noise reset of a step run on the engine's own
    # stream while the caller's stream is free to go on (the next agent's sample + forward, or the
    # priority write-back): everything that reads or writes this engine's buffers first joins.
    _bwd_stream = None
    _opt_done = None
    _readback = None

    def readback(self):
        """(loss: float, priorities: np.ndarray[B]) of the last ``rainbow_learn(host_readback=True)``."""
        host, ev, B = self._readback
        ev.synchronize()
        return float(host[B]), host[:B].numpy().copy()

    def join(self) -> None:
        """Make the current stream wait for an overlapped backward/optimiser tail, if one is pending."""
        if self._opt_done is not None:
            self._opt_done.wait()                  # current stream of the (current) device waits
            self._opt_done = None

    def workspace(self, rows: int, backward: bool) -> torch.Tensor:
        key = (rows, backward)
        ws = self._ws.get(key)
        if ws is None:
            need = ctypes.c_size_t(0)
            _lib.check(self.lib.b2rl_net_workspace_bytes(ctypes.byref(self.layout.desc), rows, int(backward),
                                                         ctypes.byref(need)))
            ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    @property
    def noise_count(self) -> int:
        return int(self._noise_count.value)

    # -- noise ---------------------------------------------------------------------------------
    def reset_noise(self, net: NetBuffers, normals: torch.Tensor | None = None) -> None:
        self.join()
        """NoisyLinear.reset_noise for every noisy layer (custom_components.py:116-131)."""
        if self.noise_count == 0:
            return
        stream = _lib.stream_ptr(self.device)
        if normals is not None:
            z = normals.to(self.device, dtype=torch.float32).contiguous()
            assert z.numel() == self.noise_count, f"need {self.noise_count} normals, got {z.numel()}"
            _lib.check(self.lib.b2rl_noise_reset_from_normals(ctypes.byref(self.layout.desc), net.eps.data_ptr(),
                                                              z.data_ptr(), stream))
            self._keep = z
        else:
            _lib.check(self.lib.b2rl_noise_reset_philox(ctypes.byref(self.layout.desc), net.eps.data_ptr(),
                                                        self.philox_seed, self.philox_offset, stream))
            self.philox_offset += self.noise_count

    # -- forward -------------------------------------------------------------------------------
    def q_values(self, net: NetBuffers, obs: torch.Tensor, support: torch.Tensor | None, use_noise: bool,
                 row_idx: torch.Tensor | None = None, want_argmax: bool = False):
        self.join()
        desc = self.layout.desc
        rows = obs.shape[0] if row_idx is None else row_idx.numel()
        obs = self._obs(obs)
        q = torch.empty((rows, desc.n_actions), dtype=torch.float32, device=self.device)
        am = torch.empty(rows, dtype=torch.int64, device=self.device) if want_argmax else None
        ws = self.workspace(rows, False)
        _lib.check(self.lib.b2rl_net_forward_q(
            ctypes.byref(desc), net.params.data_ptr(), net.eps.data_ptr(), int(use_noise),
            None if support is None else support.data_ptr(), obs.data_ptr(),
            None if row_idx is None else row_idx.data_ptr(), rows, q.data_ptr(),
            None if am is None else am.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(self.device)))
        return (q, am) if want_argmax else q

    def distributions(self, net: NetBuffers, obs: torch.Tensor, use_noise: bool, log: bool = False) -> torch.Tensor:
        """RainbowQNetwork.forward(obs, q=False, log=log): [rows, n_actions, n_atoms] (q_networks.py:265-284)."""
        self.join()
        desc = self.layout.desc
        obs = self._obs(obs)
        rows = obs.shape[0]
        out = torch.empty((rows, desc.n_actions, desc.n_atoms), dtype=torch.float32, device=self.device)
        ws = self.workspace(rows, False)
        _lib.check(self.lib.b2rl_net_forward_dist(
            ctypes.byref(desc), net.params.data_ptr(), net.eps.data_ptr(), int(use_noise), obs.data_ptr(), None, rows,
            int(log), out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(self.device)))
        return out

    def _obs(self, obs: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda_tensor(obs, "observation batch")
        want = torch.uint8 if self.layout.desc.obs_u8 else torch.float32
        if obs.dtype != want:
            obs = obs.to(want)
        return obs.contiguous()

    @staticmethod
    def _vec(t: torch.Tensor, device) -> torch.Tensor:
        if t.dtype != torch.float32 or t.device != device:
            t = t.to(device, dtype=torch.float32)
        t = t.reshape(-1)
        return t if t.is_contiguous() else t.contiguous()

    # -- learn ---------------------------------------------------------------------------------
    def _cfg(self, B, *, gamma, v_min=0.0, v_max=0.0, delta_z=1.0, weights_mode=0, driver_shapes=0, double=0,
             clip=1, lr=1e-4, tau=1e-3, prior_eps=1e-6, accumulate=0, use_noise=1, step=1,
             side_streams=0) -> _lib.LearnCfg:
        c = _lib.LearnCfg()
        c.batch = B
        c.gamma, c.v_min, c.v_max, c.delta_z = float(gamma), float(v_min), float(v_max), float(delta_z)
        c.weights_mode, c.driver_shapes, c.double_dqn, c.clip = weights_mode, driver_shapes, double, clip
        c.max_grad_norm = 10.0
        c.lr, c.beta1, c.beta2, c.adam_eps = float(lr), 0.9, 0.999, 1e-8      # torch.optim.Adam defaults
        c.bias_correction1 = 1.0 - 0.9 ** step
        c.bias_correction2 = 1.0 - 0.999 ** step
        c.tau, c.prior_eps = float(tau), float(prior_eps)
        c.accumulate, c.use_noise = accumulate, use_noise
        c.side_streams = side_streams      # bit 0: target forward, bit 1: weight gradients on library side streams
        return c

    def _bufs(self, B, batch: dict, weights, support, loss_elem, priorities, loss_scalar, proj, row_idx=None):
        b = _lib.LearnBufs()
        b.actor_params, b.target_params = self.actor.params.data_ptr(), self.target.params.data_ptr()
        b.actor_eps, b.target_eps = self.actor.eps.data_ptr(), self.target.eps.data_ptr()
        b.grads, b.exp_avg, b.exp_avg_sq = self.grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        keep = [batch["obs"] if row_idx is not None else self._obs(batch["obs"]),
                batch["next_obs"] if row_idx is not None else self._obs(batch["next_obs"]),
                self._vec(batch["action"], self.device),
                self._vec(batch["reward"], self.device), self._vec(batch["done"], self.device)]
        b.obs, b.next_obs, b.action, b.reward, b.done = (t.data_ptr() for t in keep)
        if row_idx is not None:
            b.row_idx = row_idx.data_ptr()
            keep.append(row_idx)           # read again by the (possibly overlapped) backward
        else:
            for t in keep[2:]:
                assert t.numel() == B, f"batch fields must have exactly batch_size={B} rows (quirk Q17)"
        if weights is not None:
            w = self._vec(weights, self.device)
            keep.append(w)
            b.weights = w.data_ptr()
        if support is not None:
            b.support = support.data_ptr()
        b.loss_elem = loss_elem.data_ptr()
        b.priorities = None if priorities is None else priorities.data_ptr()
        b.loss_scalar = loss_scalar.data_ptr()
        b.proj_dist = None if proj is None else proj.data_ptr()
        ws = self.workspace(B, True)
        b.workspace, b.workspace_bytes = ws.data_ptr(), ws.numel()
        return b, keep

    def rainbow_learn(self, passes: list, *, B: int, support: torch.Tensor, weights, weights_mode: int, hp: dict,
                      noise_normals=None, want_proj: bool = False, row_idx=None, overlap: bool = False,
                      after_loss=None, side_streams: int | None = None, host_readback: bool = False):
        """``passes`` = [(batch, gamma, driver_shapes)], one per ``_dqn_loss`` call of the reference
        (1-step and/or n-step; two entries when combined_reward).  Returns device tensors
        (loss_scalar[1], loss_elem[B], priorities[B], proj).

        ``after_loss(priorities)`` (optional) is called once the loss / priorities of the last pass are
        enqueued — before the backward — so the caller can enqueue the tree write-back there.
        ``overlap`` (single pass only): backward + optimiser + noise reset go to the engine's own
        stream; the caller's stream only carries sample, forward, loss and ``after_loss``.
        ``host_readback``: the scalar loss and the priorities are copied to pinned host memory as soon
        as they exist — before the backward is enqueued — and ``readback()`` waits for that copy only,
        so a caller that needs Python numbers (the reference's ``learn`` return value) does not wait
        for the rest of the step."""
        self.join()
        if (_GRAPH and host_readback and noise_normals is None and len(passes) == 1 and not want_proj and row_idx is None
                and after_loss is None and self.noise_count > 0):
            done = self._api_graph_learn(passes[0], B=B, support=support, weights=weights, weights_mode=weights_mode, hp=hp,
                                         side_streams=side_streams)
            if done:
                return None
        desc = ctypes.byref(self.layout.desc)
        loss_elem = torch.empty(B, dtype=torch.float32, device=self.device)
        out = torch.empty(B + 1, dtype=torch.float32, device=self.device)     # priorities | scalar loss (one D2H)
        priorities, loss_scalar = out[:B], out[B:]
        proj = torch.empty((B, self.layout.desc.n_atoms), dtype=torch.float32, device=self.device) if want_proj else None
        self.step += 1
        keepalive = []
        za, zt = noise_normals if noise_normals is not None else (None, None)
        overlap = overlap and len(passes) == 1
        # the target forward always forks; the weight gradients fork when the agent's tail is not already
        # running under other agents' work (then the extra streams only add contention)
        if _SIDE is not None:
            side_streams = _SIDE
        elif side_streams is None:
            side_streams = 1 if overlap else 3
        cfg = bufs = None
        for i, (batch, gamma, driver) in enumerate(passes):
            cfg = self._cfg(B, gamma=gamma, v_min=hp["v_min"], v_max=hp["v_max"], delta_z=hp["delta_z"],
                            weights_mode=weights_mode, driver_shapes=int(driver), clip=1, lr=hp["lr"], tau=hp["tau"],
                            prior_eps=hp["prior_eps"], accumulate=int(i > 0), use_noise=1, step=self.step,
                            side_streams=side_streams)
            bufs, keep = self._bufs(B, batch, weights, support, loss_elem, priorities, loss_scalar, proj, row_idx)
            keepalive.append(keep)
            _lib.check(self.lib.b2rl_rainbow_loss(desc, ctypes.byref(cfg), ctypes.byref(bufs), _lib.stream_ptr(self.device)))
            if i + 1 < len(passes):
                _lib.check(self.lib.b2rl_rainbow_backward(desc, ctypes.byref(cfg), ctypes.byref(bufs),
                                                          _lib.stream_ptr(self.device)))
        keepalive.append([loss_elem, out, proj, support, weights])
        self._keepalive = keepalive         # everything the tail reads stays allocated until the next join
        if host_readback:
            host = self._host_out.get(B)
            if host is None:
                host = self._host_out[B] = torch.empty(B + 1, dtype=torch.float32).pin_memory()
            host.copy_(out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._readback = (host, ev, B)

        def tail():
            stream = _lib.stream_ptr(self.device)
            _lib.check(self.lib.b2rl_rainbow_backward(desc, ctypes.byref(cfg), ctypes.byref(bufs), stream))
            _lib.check(self.lib.b2rl_optim_step(desc, ctypes.byref(cfg), ctypes.byref(bufs), stream))
            # dqn_rainbow.py:484-485 — actor.reset_noise() then actor_target.reset_noise()
            self.reset_noise(self.actor, za)
            self.reset_noise(self.target, zt)

        if overlap:
            fwd_done = torch.cuda.Event()
            fwd_done.record()
            if after_loss is not None:
                after_loss(priorities)
            if self._bwd_stream is None:
                self._bwd_stream = torch.cuda.Stream(device=self.device)
            with torch.cuda.stream(self._bwd_stream):
                self._bwd_stream.wait_event(fwd_done)
                tail()
                done = torch.cuda.Event()
                done.record(self._bwd_stream)
            self._opt_done = done
        else:
            if after_loss is not None:
                after_loss(priorities)
            tail()
        return loss_scalar, loss_elem, priorities, proj

    def _api_graph_learn(self, one_pass, *, B, support, weights, weights_mode, hp, side_streams) -> bool:
        """Graph replay of ``rainbow_learn`` when this exact set of device buffers has been seen before (see
        ``_ApiPlan``).  Returns False when the call has to run eagerly (first sightings, host tensors, odd dtypes)."""
        batch, gamma, driver = one_pass
        want = torch.uint8 if self.layout.desc.obs_u8 else torch.float32
        ts = [batch["obs"], batch["next_obs"], batch["action"], batch["reward"], batch["done"]]
        if weights is not None:
            ts.append(weights)
        for i, t in enumerate(ts):
            if not (isinstance(t, torch.Tensor) and t.device == self.device and t.is_contiguous()
                    and t.dtype == (want if i < 2 else torch.float32)):
                return False
        for t in ts[2:5]:
            if t.numel() != B:
                return False
        if ts[0].numel() != B * self.layout.desc.obs_elems or ts[1].numel() != ts[0].numel():
            return False
        if weights is not None and weights.numel() != B:
            return False
        if _SIDE is not None:
            side_streams = _SIDE
        elif side_streams is None:
            side_streams = 1
        key = (B, weights_mode, side_streams, float(gamma), bool(driver), support.data_ptr(), float(hp["v_min"]),
               float(hp["v_max"]), float(hp["delta_z"]), float(hp["tau"]), float(hp["prior_eps"]),
               *(t.data_ptr() for t in ts))
        plan = self._api_plans.get(key)
        if plan is None:
            if len(self._api_plans) >= 16:                   # addresses keep changing: stop trying to cache them
                return False
            self._api_plans[key] = _ApiPlan(self, B)
            return False
        plan.seen += 1
        st = plan.state_host
        self.step += 1
        st.lr = float(hp["lr"])
        st.bias_correction1 = 1.0 - 0.9 ** self.step
        st.bias_correction2 = 1.0 - 0.999 ** self.step
        st.noise_offset[0] = self.philox_offset
        st.noise_offset[1] = self.philox_offset + self.noise_count
        self.philox_offset += 2 * self.noise_count
        if plan.front is None:
            plan.capture(batch, gamma, driver, weights, weights_mode, support, hp, side_streams)
        lib = self.lib
        _lib.check(lib.b2rl_graph_launch(plan.front, ctypes.byref(st), _lib.stream_ptr(self.device)))
        plan.rb_ev.record()
        self._readback = (plan.host, plan.rb_ev, B)
        # backward + optimiser + noise reset under the caller's next calls (sampling, ingest, the next agent)
        if self._bwd_stream is None:
            self._bwd_stream = torch.cuda.Stream(device=self.device)
        self._bwd_stream.wait_event(plan.rb_ev)
        _lib.check(lib.b2rl_graph_launch(plan.tail, None, self._bwd_stream.cuda_stream))
        plan.done.record(self._bwd_stream)
        self._opt_done = plan.done
        self._keepalive = (ts, support)          # the tail reads the batch: keep it allocated until the next join
        return True

    def dqn_learn(self, batch: dict, *, B: int, hp: dict, double: bool):
        self.join()
        desc = ctypes.byref(self.layout.desc)
        loss_elem = torch.empty(B, dtype=torch.float32, device=self.device)
        loss_scalar = torch.empty(1, dtype=torch.float32, device=self.device)
        self.step += 1
        cfg = self._cfg(B, gamma=hp["gamma"], double=int(double), clip=0, lr=hp["lr"], tau=hp["tau"], use_noise=0,
                        step=self.step)
        bufs, keep = self._bufs(B, batch, None, None, loss_elem, None, loss_scalar, None)
        _lib.check(self.lib.b2rl_dqn_learn(desc, ctypes.byref(cfg), ctypes.byref(bufs), _lib.stream_ptr(self.device)))
        self._keepalive = keep
        return loss_scalar

    # -- fused HBM-resident step, CUDA-graph replay ----------------------------------------------
    def _plan_key(self, per, nmem, B, support, hp, gamma_n, weights_mode, side_streams):
        f = nmem._fields
        return (id(per), id(nmem), B, weights_mode, side_streams, float(gamma_n), support.data_ptr(),
                float(hp["v_min"]), float(hp["v_max"]), float(hp["delta_z"]), float(hp["tau"]), float(hp["prior_eps"]),
                per.sum_tree.data_ptr, per.min_tree.data_ptr, per._cap, float(per.alpha), per._philox_seed,
                f[("obs",)].data_ptr(), f[(nmem.ns_key,)].data_ptr(), f[("action",)].data_ptr())

    def _fused_graph_step(self, per, nmem, *, B, beta, support, hp, gamma_n, weights_mode, overlap, side_streams):
        """``rainbow_fused_step`` with the whole chain replayed from two CUDA graphs per (agent, replay pair, B):
        *front* = step-state write -> PER sample -> 3 forwards -> projection / loss -> priority write-back (the part
        the next agent's sampler depends on) and *tail* = backward -> clip/Adam/Polyak -> noise reset x2.  The first
        call of a plan runs eagerly (it also creates the library's side streams and kernel attributes), the second
        captures, every later one replays.  The replay reads this step's scalars (beta, len(memory), Philox
        offsets, Adam's lr and bias corrections) from a device block its first node rewrites — same arithmetic,
        same random streams and bit-identical results as the eager path (tests/test_graph_gpu.py)."""
        key = self._plan_key(per, nmem, B, support, hp, gamma_n, weights_mode, side_streams)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = _FusedPlan(self, per, nmem, B, support, hp, gamma_n, weights_mode, side_streams)
            return None                                   # caller runs this step eagerly
        st = plan.state_host
        st.beta, st.size = float(beta), int(per._size)
        st.sample_offset = per._philox_offset
        st.noise_offset[0] = self.philox_offset
        st.noise_offset[1] = self.philox_offset + self.noise_count
        self.step += 1
        st.lr = float(hp["lr"])
        st.bias_correction1 = 1.0 - 0.9 ** self.step
        st.bias_correction2 = 1.0 - 0.999 ** self.step
        per._philox_offset += B
        self.philox_offset += 2 * self.noise_count
        if plan.front is None:
            plan.capture()
        lib, cur = self.lib, _lib.stream_ptr(self.device)
        per._own_max_on_device()               # the captured write-back folds max(priority) into the device scalar
        _lib.check(lib.b2rl_graph_launch(plan.front, ctypes.byref(st), cur))
        if overlap:
            fwd_done = plan.fwd_done
            fwd_done.record()
            if self._bwd_stream is None:
                self._bwd_stream = torch.cuda.Stream(device=self.device)
            self._bwd_stream.wait_event(fwd_done)
            _lib.check(lib.b2rl_graph_launch(plan.tail, None, self._bwd_stream.cuda_stream))
            plan.done.record(self._bwd_stream)
            self._opt_done = plan.done
        else:
            _lib.check(lib.b2rl_graph_launch(plan.tail, None, cur))
        return plan.out[B:], plan.idx, plan.out[:B]

    # -- fused HBM-resident step ---------------------------------------------------------------
    def rainbow_fused_step(self, per, n_step_memory, *, B: int, beta: float, support: torch.Tensor, hp: dict,
                           gamma_n: float, weights_mode: int = 1, uniforms=None, noise_normals=None,
                           overlap: bool = False, side_streams: int | None = None, graph: bool | None = None):
        """One gradient step with the replay resident in HBM and no host round trip:
        sample (tree descent + IS weights + n-step scalars, one kernel) -> learn with the encoder
        reading frames from the ring through the sampled indices -> priorities written back into
        the trees on device.  Equivalent to train_off_policy.py:399-412 for one agent.

        The priorities exist as soon as the loss does, so the tree write-back is enqueued before the
        backward: with ``overlap`` the next sampler on this stream (the next agent of the population
        sharing the buffer) sees exactly the tree the sequential loop would show it, while this
        agent's backward + optimiser still run on its own stream."""
        self.join()
        if graph is None:
            graph = _GRAPH
        if graph and uniforms is None and noise_normals is None and per.device_rng and self.noise_count > 0:
            if _SIDE is not None:
                side_streams = _SIDE
            elif side_streams is None:
                side_streams = 1 if overlap else 3
            out = self._fused_graph_step(per, n_step_memory, B=B, beta=beta, support=support, hp=hp, gamma_n=gamma_n,
                                         weights_mode=weights_mode, overlap=overlap, side_streams=side_streams)
            if out is not None:
                return out
        idx, w, a, r, d = per.sample_fused(B, beta, n_step_memory, uniforms)
        f = n_step_memory._fields
        batch = dict(obs=f[("obs",)], next_obs=f[(n_step_memory.ns_key,)], action=a, reward=r, done=d)
        loss, loss_elem, pri, _ = self.rainbow_learn([(batch, gamma_n, False)], B=B, support=support, weights=w,
                                                     weights_mode=weights_mode, hp=hp, noise_normals=noise_normals,
                                                     row_idx=idx, overlap=overlap, side_streams=side_streams,
                                                     after_loss=lambda p: per.update_priorities_device(idx, p))
        return loss, idx, pri
