"""``TD3`` and ``DDPG`` — drop-ins for agilerl/algorithms/td3.py:30-640 and ddpg.py:30-600 on the CUDA path
(BASELINE configs[2]: 17-dim observations, 6-dim actions, batch 512).

Same constructors (td3.py:93-120 / ddpg.py:90-116), attributes (``actor, actor_target, critic_1, critic_2,
critic_target_1, critic_target_2`` / ``critic, critic_target``; ``learn_counter, policy_freq, current_noise`` ...),
``learn(experiences, noise_clip=0.5, policy_noise=0.2) -> (actor_loss | None, critic_loss)`` (td3.py:462-551,
ddpg.py:422-494), ``get_action`` with Ornstein-Uhlenbeck / Gaussian exploration noise from NumPy's global stream
(td3.py:392-452), ``soft_update``, ``test``.  The body of ``learn`` — 5 network forwards, TD target, MSE, critic
backward + Adam, the every-``policy_freq`` actor step through the updated critic_1 and the Polyak updates — is
``b2rl_ddpg_learn`` (csrc/ddpg.cuh): 23 launches on a TD3 policy step.

Kept quirk: the target-policy noise is drawn IN PLACE on the batch's ``action`` tensor
(``actions.data.normal_(0, policy_noise)``, td3.py:497): after ``learn`` the caller's tensor holds the noise.
"""
from __future__ import annotations

import copy
import ctypes
import warnings
from typing import Any

import numpy as np
import torch

from .. import _lib
from ..compat import spaces
from ..networks.actors import ContinuousQNetwork, DeterministicActor
from .core.base import RLAlgorithm
from .core.registry import HyperparameterConfig, NetworkGroup, OptimizerConfig


class _AdamState:
    """Adam moments + step count of one flat parameter buffer (what ``OptimizerWrapper`` exposes in the reference)."""

    def __init__(self, net, lr: float):
        n = max(net.layout.n_params, 1)
        dev = net.buffers.params.device
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step = 0
        self.lr = lr

    def state_dict(self) -> dict:
        return {"step": self.step, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "lr": self.lr}

    def load_state_dict(self, sd: dict, strict: bool = False) -> None:
        if sd["exp_avg"].numel() != self.exp_avg.numel():
            if strict:
                raise ValueError("optimizer state does not fit this network")
            warnings.warn("optimizer state does not fit this network: Adam moments start fresh", stacklevel=2)
            return
        self.step = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"].to(self.exp_avg.device))
        self.exp_avg_sq.copy_(sd["exp_avg_sq"].to(self.exp_avg.device))

    def zero_grad(self) -> None:
        pass

    def __repr__(self) -> str:                    # what ``str(torch.optim.Adam)`` shows of the hyper-parameters
        return f"Adam (lr: {self.lr}, betas: (0.9, 0.999), eps: 1e-08, weight_decay: 0)"


class _DeterministicPG(RLAlgorithm):
    _twin = False
    _name = "DDPG"
    _clone_skip = ("_ws", "_keep", "_low_dev", "_high_dev", "actor_optimizer", "critic_optimizer", "critic_1_optimizer",
                   "critic_2_optimizer", "action_low", "action_high")

    def __init__(self, observation_space, action_space, O_U_noise: bool = True, vect_noise_dim: int = 1,
                 expl_noise: float = 0.1, mean_noise: float = 0.0, theta: float = 0.15, dt: float = 1e-2, index: int = 0,
                 hp_config: HyperparameterConfig | None = None, net_config: dict | None = None, batch_size: int = 64,
                 lr_actor: float = 1e-4, lr_critic: float = 1e-3, learn_step: int = 5, gamma: float = 0.99, tau: float = 1e-3,
                 normalize_images: bool = True, mut: str | None = None, policy_freq: int = 2, actor_network=None,
                 critic_networks=None, share_encoders: bool = False, device: str = "cuda", accelerator: Any | None = None,
                 wrap: bool = True) -> None:
        super().__init__(observation_space, action_space, index=index, hp_config=hp_config, device=device,
                         accelerator=accelerator, normalize_images=normalize_images, name=self._name)
        assert isinstance(action_space, spaces.Box), f"{self._name} only supports continuous action spaces."
        self.action_dim = int(action_space.shape[0])
        assert isinstance(expl_noise, (float, int)) or (
            isinstance(expl_noise, np.ndarray) and expl_noise.shape == (vect_noise_dim, self.action_dim)), (
            "Exploration action noise rate must be a float, or an array of size action_dim")
        if isinstance(expl_noise, (float, int)):
            assert expl_noise >= 0, "Exploration noise must be greater than or equal to zero."
        assert learn_step >= 1, "Learn step must be greater than or equal to one."
        assert isinstance(learn_step, int), "Learn step rate must be an integer."
        assert isinstance(batch_size, int), "Batch size must be an integer."
        assert batch_size >= 1, "Batch size must be greater than or equal to one."
        assert isinstance(lr_actor, float), "Actor learning rate must be a float."
        assert lr_actor > 0, "Actor learning rate must be greater than zero."
        assert isinstance(lr_critic, float), "Critic learning rate must be a float."
        assert lr_critic > 0, "Critic learning rate must be greater than zero."
        assert isinstance(gamma, (float, int, torch.Tensor)), "Gamma must be a float."
        assert isinstance(tau, float), "Tau must be a float."
        assert tau > 0, "Tau must be greater than zero."
        assert isinstance(policy_freq, int), "Policy frequency must be an integer."
        assert policy_freq >= 1, "Policy frequency must be greater than or equal to one."
        assert isinstance(wrap, bool), "Wrap models flag must be boolean value True or False."
        self.batch_size, self.lr_actor, self.lr_critic, self.learn_step = batch_size, lr_actor, lr_critic, learn_step
        self.gamma, self.tau, self.mut, self.policy_freq, self.net_config = gamma, tau, mut, policy_freq, net_config
        self.O_U_noise, self.vect_noise_dim, self.share_encoders = O_U_noise, vect_noise_dim, share_encoders
        self.current_noise = np.zeros((vect_noise_dim, self.action_dim))
        self.theta, self.dt, self.learn_counter = theta, dt, 0
        self.action_low = torch.as_tensor(action_space.low, dtype=torch.float32)
        self.action_high = torch.as_tensor(action_space.high, dtype=torch.float32)
        self.expl_noise = expl_noise if isinstance(expl_noise, np.ndarray) else expl_noise * np.ones((vect_noise_dim, self.action_dim))
        self.mean_noise = mean_noise if isinstance(mean_noise, np.ndarray) else mean_noise * np.ones((vect_noise_dim, self.action_dim))

        n_c = 2 if self._twin else 1
        if actor_network is not None and critic_networks is not None:
            if self._twin:
                assert isinstance(critic_networks, (list, tuple)), "Critic network must be a list or tuple"
                assert len(critic_networks) == 2, "TD3 requires exactly 2 critic networks."
            else:
                critic_networks = critic_networks if isinstance(critic_networks, (list, tuple)) else [critic_networks]
            if not isinstance(actor_network, DeterministicActor):
                raise TypeError(f"Passed actor network is of type {type(actor_network)}, but must be of type EvolvableModule.")
            for i, c in enumerate(critic_networks):
                if not isinstance(c, ContinuousQNetwork):
                    raise TypeError(f"Passed critic network at index {i} is of type {type(c)}, but must be of type EvolvableModule.")
            self.actor, self.actor_target = actor_network.clone(), actor_network.clone()
            critics = [c.clone() for c in critic_networks[:n_c]]
            targets = [c.clone() for c in critic_networks[:n_c]]
        else:
            if (actor_network is not None) != (critic_networks is not None):
                warnings.warn("Actor and critic networks must both be supplied to use custom networks. Defaulting to net config.",
                              stacklevel=2)
            net_config = {} if net_config is None else copy.deepcopy(net_config)
            encoder_config = net_config.get("encoder_config", None)
            if encoder_config is not None:
                if encoder_config.get("layer_norm", False):
                    warnings.warn(f"Layer normalization is not supported for the encoder of {self._name} networks. Disabling it.",
                                  stacklevel=2)
                encoder_config["layer_norm"] = False
            else:
                encoder_config = dict(hidden_size=[64, 64], output_activation="ReLU", layer_norm=False, output_vanish=False)
            net_config["encoder_config"] = encoder_config
            head_config = net_config.get("head_config", None)
            critic_head = copy.deepcopy(head_config) if head_config is not None else dict(hidden_size=[64])
            critic_head["output_activation"] = None
            critic_cfg = copy.deepcopy(net_config)
            critic_cfg["head_config"] = critic_head
            self.net_config = net_config
            mk_a = lambda: DeterministicActor(observation_space, action_space, device=self.device, **net_config)
            mk_c = lambda: ContinuousQNetwork(observation_space, action_space, device=self.device, **critic_cfg)
            self.actor, self.actor_target = mk_a(), mk_a()
            critics, targets = [mk_c() for _ in range(n_c)], [mk_c() for _ in range(n_c)]
        self._set_critics(critics, targets)
        if self.share_encoders:                                    # ddpg.py:289-296 / td3.py: before the targets are initialised
            self.share_encoder_parameters()
            self.register_mutation_hook(self.share_encoder_parameters)
        self.actor_target.load_state_dict(self.actor.state_dict())
        for c, t in zip(critics, targets):
            t.load_state_dict(c.state_dict())
        self.register_network_group(NetworkGroup(eval_network="actor", shared_networks="actor_target", policy=True))
        for cn, tn in zip(self._critic_names, self._target_names):
            self.register_network_group(NetworkGroup(eval_network=cn, shared_networks=tn))
        self.registry.register_optimizer(OptimizerConfig(name="actor_optimizer", networks=["actor"], lr="lr_actor"))
        for i, cn in enumerate(self._critic_names):
            self.registry.register_optimizer(OptimizerConfig(name=self._opt_names[i], networks=[cn], lr="lr_critic"))
        self._bind_engine()

    # -- network bookkeeping --------------------------------------------------------------------------
    def _set_critics(self, critics, targets) -> None:
        for name, net in zip(self._critic_names + self._target_names, list(critics) + list(targets)):
            setattr(self, name, net)

    def _critics(self):
        return [getattr(self, n) for n in self._critic_names]

    def share_encoder_parameters(self) -> None:
        """ddpg.py:335-351 -> utils/algo_utils.py:161-184: the critics' (and target critics') ENCODER parameters become
        detached copies of the actor's encoder — a snapshot taken here (construction, and again after every mutation through
        the mutation hook) that no optimiser touches from then on.  The fused learn call updates every critic parameter, so
        ``learn`` saves the critics' encoder block before the call and puts it back afterwards: identical to never having
        stepped it (the critics' heads and the actor never read the encoder's own update), and the frozen values live in
        the critics' own buffers, so clones, checkpoints and cross-rank moves carry them."""
        src = self.actor
        keys = [k for k, e in src.layout.entries.items() if k.startswith("encoder.") and e.buf == "param"]
        for net in self._critics() + self._targets():
            for k in keys:
                if k not in net.layout.entries or net.layout.entries[k].shape != src.layout.entries[k].shape:
                    raise KeyError(f"Found incompatible encoder architectures: {k} not found in shared network.")
            for k in keys:
                net.buffers.view(k).copy_(src.buffers.view(k))

    @staticmethod
    def _encoder_block(net) -> int:
        """Length of the leading block of the flat parameter buffer that holds the encoder (FlatLayout lays the encoder out
        first): everything before the first head parameter."""
        offs = [e.offset for k, e in net.layout.entries.items() if e.buf == "param" and not k.startswith("encoder.")]
        return min(offs) if offs else net.layout.n_params

    def _targets(self):
        return [getattr(self, n) for n in self._target_names]

    def _bind_engine(self, keep: dict | None = None) -> None:
        self.actor_optimizer = _AdamState(self.actor, self.lr_actor)
        for name, c in zip(self._opt_names, self._critics()):
            setattr(self, name, _AdamState(c, self.lr_critic))
        if keep:
            self.actor_optimizer.load_state_dict(keep["actor"])
            for name, sd in zip(self._opt_names, keep["critics"]):
                getattr(self, name).load_state_dict(sd)
        self._ws: dict = {}
        self._low_dev = self.action_low.to(self._dev)
        self._high_dev = self.action_high.to(self._dev)
        self._noise_offset = 0
        self.optimizer = self.actor_optimizer            # single-optimiser hooks of the base class (checkpoints)

    def _opt_state(self) -> dict:
        return {"actor": self.actor_optimizer.state_dict(), "critics": [getattr(self, n).state_dict() for n in self._opt_names]}

    def reinit_optimizers(self, optimizer=None) -> None:
        self._bind_engine()

    # -- cross-rank move (population sharding: hpo/tournament.py::_select_sharded broadcasts a winner from its owner) ------
    def _all_optimizers(self) -> list:
        return [self.actor_optimizer] + [getattr(self, n) for n in self._opt_names]

    def export_state(self):
        """-> (picklable description, [device tensors]): every network's flat parameter buffer, then both Adam moments of
        the actor's and of every critic's optimiser."""
        nets = self._evolvable_attrs()
        meta = {"init": {k: v for k, v in self._init_kwargs().items() if k not in ("device", "accelerator")},
                "nets": {n: getattr(self, n).init_dict for n in nets},
                "attrs": {"scores": list(self.scores), "fitness": list(self.fitness), "steps": list(self.steps), "mut": self.mut,
                          "index": self.index, "learn_counter": self.learn_counter,
                          "opt_steps": [o.step for o in self._all_optimizers()], "current_noise": self.current_noise}}
        for n in nets:
            meta["nets"][n].pop("device", None)
        tensors = [getattr(self, n).buffers.params for n in nets]
        for o in self._all_optimizers():
            tensors += [o.exp_avg, o.exp_avg_sq]
        return meta, tensors

    @classmethod
    def from_state(cls, meta, tensors, like):
        agent = cls(device=like.device, **meta["init"])
        nets = agent._evolvable_attrs()
        for n in nets:                                  # the moved member's (possibly mutated) architectures
            kw = dict(meta["nets"][n])
            kw["device"] = like.device
            setattr(agent, n, type(getattr(agent, n))(**kw))
        agent._after_network_swap()
        it = iter(tensors)
        for n in nets:
            getattr(agent, n).buffers.params.copy_(next(it))
        a = meta["attrs"]
        for o, step in zip(agent._all_optimizers(), a["opt_steps"]):
            o.exp_avg.copy_(next(it)); o.exp_avg_sq.copy_(next(it))
            o.step = step
        agent.scores, agent.fitness, agent.steps, agent.mut = a["scores"], a["fitness"], a["steps"], a["mut"]
        agent.index, agent.learn_counter, agent.current_noise = a["index"], a["learn_counter"], a["current_noise"]
        return agent

    # -- checkpoints: the base class knows ONE optimiser; these learners own two or three --------------------------------
    def save_checkpoint(self, path: str) -> None:
        super().save_checkpoint(path)
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        ckpt["optimizers"] = self._opt_state()
        ckpt["learn_counter"] = self.learn_counter
        torch.save(ckpt, path)

    def load_checkpoint(self, path: str) -> None:
        super().load_checkpoint(path)
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        if "optimizers" in ckpt:
            self.actor_optimizer.load_state_dict(ckpt["optimizers"]["actor"], strict=True)
            for name, sd in zip(self._opt_names, ckpt["optimizers"]["critics"]):
                getattr(self, name).load_state_dict(sd, strict=True)
        self.learn_counter = ckpt.get("learn_counter", self.learn_counter)

    def _after_network_swap(self) -> None:
        self._bind_engine()

    def _copy_networks_to(self, clone) -> None:
        clone.actor, clone.actor_target = self.actor.clone(), self.actor_target.clone()
        clone._set_critics([c.clone() for c in self._critics()], [t.clone() for t in self._targets()])
        clone._bind_engine(keep=self._opt_state())

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name == "lr_actor" and "actor_optimizer" in self.__dict__:
            self.actor_optimizer.lr = value
        if name == "lr_critic":
            for n in getattr(self, "_opt_names", ()):
                if n in self.__dict__:
                    getattr(self, n).lr = value

    # -- acting (td3.py:392-452) --------------------------------------------------------------------------
    def get_action(self, obs, training: bool = True, *args, **kwargs) -> np.ndarray:
        obs = self.preprocess_observation(obs)
        action = self.actor(obs)
        if training:
            action = action.cpu().numpy()
            return (action + self.action_noise()).clip(-1, 1)
        action = DeterministicActor.rescale_action(action.cpu(), self.action_low, self.action_high, self.actor.output_activation)
        return action.numpy()

    def action_noise(self) -> np.ndarray:
        if self.O_U_noise:
            noise = (self.current_noise + self.theta * (self.mean_noise - self.current_noise) * self.dt
                     + self.expl_noise * np.sqrt(self.dt) * np.random.normal(size=(self.vect_noise_dim, self.action_dim)))
            self.current_noise = noise
        else:
            noise = np.random.normal(self.mean_noise, self.expl_noise, size=(self.vect_noise_dim, self.action_dim))
        return noise.astype(np.float32)

    def reset_action_noise(self, indices) -> None:
        self.current_noise[indices] = self.mean_noise[indices]

    # -- learning -----------------------------------------------------------------------------------------
    def _workspace(self, B: int) -> torch.Tensor:
        ws = self._ws.get(B)
        if ws is None:
            need = ctypes.c_size_t(0)
            _lib.check(_lib.load().b2rl_ddpg_workspace_bytes(ctypes.byref(self.actor.layout.desc),
                                                             ctypes.byref(self._critics()[0].layout.desc), B, ctypes.byref(need)))
            ws = self._ws[B] = torch.empty(need.value, dtype=torch.uint8, device=self._dev)
        return ws

    def learn(self, experiences, noise_clip: float = 0.5, policy_noise: float = 0.2, noise: torch.Tensor | None = None):
        """td3.py:462-551 / ddpg.py:422-494.  ``noise`` (optional, [B, act_dim]) injects the N(0, policy_noise) draws
        instead of the device Philox stream (parity tests).  Returns ``(actor_loss | None, critic_loss)``."""
        out, policy_update = self.learn_device(experiences, noise_clip, policy_noise, noise)
        host = out.tolist()                                   # the reference returns Python floats (.item())
        return (host[1] if policy_update else None), host[0]

    def learn_device(self, experiences, noise_clip: float = 0.5, policy_noise: float = 0.2, noise: torch.Tensor | None = None):
        """``learn`` without the host read-back: returns (device tensor [critic_loss, actor_loss], policy_update)."""
        lib = _lib.load()
        f32 = lambda t: t if (t.dtype == torch.float32 and t.device == self._dev and t.is_contiguous()) else \
            t.to(self._dev, dtype=torch.float32).contiguous()
        obs, next_obs = f32(self.preprocess_observation(experiences["obs"])), f32(self.preprocess_observation(experiences["next_obs"]))
        action = experiences["action"]
        if not (isinstance(action, torch.Tensor) and action.dtype == torch.float32 and action.device == self._dev
                and action.is_contiguous()):
            raise _lib.B2RLError("experiences['action'] must be a contiguous float32 CUDA tensor: learn() overwrites it in "
                                 "place with the target-policy noise, like the reference (td3.py:497)")
        reward, done = f32(experiences["reward"]).reshape(-1), f32(experiences["done"]).reshape(-1)
        B = obs.shape[0]
        assert action.numel() == B * self.action_dim and reward.numel() == B and done.numel() == B
        self.learn_counter += 1
        policy_update = self.learn_counter % self.policy_freq == 0
        critics, targets = self._critics(), self._targets()
        opts = [getattr(self, n) for n in self._opt_names]
        cfg = _lib.DdpgCfg()
        cfg.batch, cfg.twin, cfg.policy_update = B, int(self._twin), int(policy_update)
        cfg.gamma, cfg.tau, cfg.noise_clip, cfg.policy_noise = float(self.gamma), float(self.tau), float(noise_clip), float(policy_noise)
        cfg.lr_actor, cfg.lr_critic, cfg.beta1, cfg.beta2, cfg.adam_eps = float(self.lr_actor), float(self.lr_critic), 0.9, 0.999, 1e-8
        for o in opts:
            o.step += 1
        cfg.bc1_critic, cfg.bc2_critic = 1.0 - 0.9 ** opts[0].step, 1.0 - 0.999 ** opts[0].step
        if policy_update:
            self.actor_optimizer.step += 1
        a_step = max(self.actor_optimizer.step, 1)
        cfg.bc1_actor, cfg.bc2_actor = 1.0 - 0.9 ** a_step, 1.0 - 0.999 ** a_step
        cfg.noise_seed, cfg.noise_offset = 0x7D3 + 104729 * int(self.index), self._noise_offset
        bufs = _lib.DdpgBufs()
        bufs.actor, bufs.actor_target = self.actor.buffers.params.data_ptr(), self.actor_target.buffers.params.data_ptr()
        ao = self.actor_optimizer
        bufs.actor_grads, bufs.actor_m, bufs.actor_v = ao.grads.data_ptr(), ao.exp_avg.data_ptr(), ao.exp_avg_sq.data_ptr()
        for i, (c, t, o) in enumerate(zip(critics, targets, opts)):
            bufs.critic[i], bufs.critic_target[i] = c.buffers.params.data_ptr(), t.buffers.params.data_ptr()
            bufs.critic_grads[i], bufs.critic_m[i], bufs.critic_v[i] = o.grads.data_ptr(), o.exp_avg.data_ptr(), o.exp_avg_sq.data_ptr()
        bufs.obs, bufs.next_obs, bufs.action = obs.data_ptr(), next_obs.data_ptr(), action.data_ptr()
        bufs.reward, bufs.done = reward.data_ptr(), done.data_ptr()
        nz = None
        if noise is not None:
            nz = f32(noise)
            assert nz.numel() == B * self.action_dim
            bufs.noise = nz.data_ptr()
        else:
            self._noise_offset += B * self.action_dim
        bufs.action_low, bufs.action_high = self._low_dev.data_ptr(), self._high_dev.data_ptr()
        out = torch.empty(2, dtype=torch.float32, device=self._dev)
        bufs.critic_loss, bufs.actor_loss = out[0:].data_ptr(), out[1:].data_ptr()
        ws = self._workspace(B)
        bufs.workspace, bufs.workspace_bytes = ws.data_ptr(), ws.numel()
        frozen = critics[0].buffers.params[:self._encoder_block(critics[0])].clone() if self.share_encoders else None
        _lib.check(lib.b2rl_ddpg_learn(ctypes.byref(self.actor.layout.desc), ctypes.byref(critics[0].layout.desc),
                                       ctypes.byref(cfg), ctypes.byref(bufs), _lib.stream_ptr(self._dev)))
        self._keep = (obs, next_obs, reward, done, nz, out)
        if frozen is not None:                                    # share_encoders: the critics' encoders stay what they were
            for net in critics + targets:
                net.buffers.params[:frozen.numel()].copy_(frozen)
        return out, policy_update

    def soft_update(self, net, target) -> None:
        """td3.py:553-565."""
        p, t = net.buffers.params, target.buffers.params
        t.copy_(self.tau * p + (1.0 - self.tau) * t)

    def test(self, env, swap_channels: bool = False, max_steps: int | None = None, loop: int = 3) -> float:
        """td3.py:567-620."""
        self.set_training_mode(False)
        rewards = []
        num_envs = env.num_envs if hasattr(env, "num_envs") else 1
        for _ in range(loop):
            obs, _ = env.reset()
            scores, completed, finished = np.zeros(num_envs), np.zeros(num_envs), np.zeros(num_envs)
            step = 0
            while not np.all(finished):
                action = self.get_action(obs, training=False)
                obs, reward, done, trunc, _ = env.step(action)
                step += 1
                scores += np.array(reward)
                for i, (d, t) in enumerate(zip(np.atleast_1d(done), np.atleast_1d(trunc))):
                    if (d or t or (max_steps is not None and step == max_steps)) and not finished[i]:
                        completed[i] = scores[i]
                        finished[i] = 1
            rewards.append(np.mean(completed))
        mean_fit = float(np.mean(rewards))
        self.fitness.append(mean_fit)
        return mean_fit


class TD3(_DeterministicPG):
    _twin, _name = True, "TD3"
    _critic_names, _target_names = ["critic_1", "critic_2"], ["critic_target_1", "critic_target_2"]
    _opt_names = ["critic_1_optimizer", "critic_2_optimizer"]

    def __init__(self, observation_space, action_space, O_U_noise: bool = True, vect_noise_dim: int = 1,
                 expl_noise: float = 0.1, mean_noise: float = 0.0, theta: float = 0.15, dt: float = 1e-2, index: int = 0,
                 hp_config: HyperparameterConfig | None = None, net_config: dict | None = None, batch_size: int = 64,
                 lr_actor: float = 1e-4, lr_critic: float = 1e-3, learn_step: int = 5, gamma: float = 0.99, tau: float = 0.005,
                 normalize_images: bool = True, mut: str | None = None, policy_freq: int = 2, actor_network=None,
                 critic_networks=None, share_encoders: bool = False, device: str = "cuda", accelerator: Any | None = None,
                 wrap: bool = True) -> None:
        super().__init__(observation_space, action_space, O_U_noise, vect_noise_dim, expl_noise, mean_noise, theta, dt, index,
                         hp_config, net_config, batch_size, lr_actor, lr_critic, learn_step, gamma, tau, normalize_images, mut,
                         policy_freq, actor_network, critic_networks, share_encoders, device, accelerator, wrap)


class DDPG(_DeterministicPG):
    _twin, _name = False, "DDPG"
    _critic_names, _target_names = ["critic"], ["critic_target"]
    _opt_names = ["critic_optimizer"]

    def __init__(self, observation_space, action_space, O_U_noise: bool = True, expl_noise: float = 0.1,
                 vect_noise_dim: int = 1, mean_noise: float = 0.0, theta: float = 0.15, dt: float = 1e-2, index: int = 0,
                 hp_config: HyperparameterConfig | None = None, net_config: dict | None = None, batch_size: int = 64,
                 lr_actor: float = 1e-4, lr_critic: float = 1e-3, learn_step: int = 5, gamma: float = 0.99, tau: float = 1e-3,
                 normalize_images: bool = True, mut: str | None = None, policy_freq: int = 2, actor_network=None,
                 critic_network=None, share_encoders: bool = False, device: str = "cuda", accelerator: Any | None = None,
                 wrap: bool = True) -> None:
        super().__init__(observation_space, action_space, O_U_noise, vect_noise_dim, expl_noise, mean_noise, theta, dt, index,
                         hp_config, net_config, batch_size, lr_actor, lr_critic, learn_step, gamma, tau, normalize_images, mut,
                         policy_freq, actor_network, None if critic_network is None else [critic_network], share_encoders, device,
                         accelerator, wrap)
