"""``DeterministicActor`` (agilerl/networks/actors.py:78-210) and ``ContinuousQNetwork``
(agilerl/networks/q_networks.py:302-443) — the networks of DDPG / TD3 — over flat HBM parameter buffers.

Both are an MLP encoder followed by a LayerNorm MLP head; the critic concatenates the action to the latent
(``torch.cat([latent, actions], dim=-1)``, q_networks.py:424-425) and its encoder drops LayerNorm
(q_networks.py:349-367).  State-dict keys are the reference's (``encoder.model.encoder_linear_layer_1.weight``,
``head_net.model.actor_linear_layer_output.bias``, ``head_net.model.value_layer_norm_1.weight`` ...).  The forward /
backward kernels are the fused chain kernels of csrc/ddpg.cuh."""
from __future__ import annotations

import copy
import ctypes

import numpy as np
import torch

from .. import _lib
from ..compat import spaces
from ..modules.mlp import EvolvableMLP
from .base import EvolvableNetwork
from .q_networks import _head_kwargs
from .spec import MlpSpec, NetSpec


def _mlp_only(observation_space) -> None:
    shape = getattr(observation_space, "shape", None)
    if not isinstance(observation_space, spaces.Box) or shape is None or len(shape) != 1:
        raise NotImplementedError("DDPG / TD3 on the CUDA path take vector observations (BASELINE configs[2]: 17-dim)")


class DeterministicActor(EvolvableNetwork):
    kind = "q"
    _allowed_output_activations = ["Tanh", "Sigmoid", "Softmax", "GumbelSoftmax"]

    def __init__(self, observation_space, action_space, encoder_cls=None, encoder_config: dict | None = None,
                 head_config: dict | None = None, min_latent_dim: int = 8, max_latent_dim: int = 128, latent_dim: int = 32,
                 simba: bool = False, recurrent: bool = False, device: str = "cuda", random_seed: int | None = None,
                 encoder_name: str = "encoder") -> None:
        _mlp_only(observation_space)
        if not isinstance(action_space, spaces.Box):
            raise NotImplementedError("the CUDA actor implements continuous (Box) action spaces: Tanh output")
        super().__init__(observation_space, encoder_cls=encoder_cls, encoder_config=encoder_config,
                         encoder_name=encoder_name, action_space=action_space, min_latent_dim=min_latent_dim,
                         max_latent_dim=max_latent_dim, latent_dim=latent_dim, simba=simba, recurrent=recurrent,
                         device=device, random_seed=random_seed)
        self.action_low = torch.as_tensor(action_space.low, dtype=torch.float32)
        self.action_high = torch.as_tensor(action_space.high, dtype=torch.float32)
        output_activation = "Tanh"
        head_config = dict(head_config) if head_config is not None else dict(hidden_size=[32])
        if head_config.get("output_activation") not in (None, "Tanh"):
            raise NotImplementedError("only the Tanh output activation is implemented in the CUDA actor")
        head_config["output_activation"] = output_activation
        self.output_activation = output_activation
        self.output_size = int(action_space.shape[0])
        self.num_actions = self.output_size
        self.head_net = EvolvableMLP(num_inputs=self.latent_dim, num_outputs=self.output_size, name="actor", device=device,
                                     random_seed=random_seed, **_head_kwargs(head_config, EvolvableMLP))
        self._finish_init()

    def _net_spec(self) -> NetSpec:
        h = self.head_net
        head = MlpSpec("head_net.model.", "actor", self.latent_dim, self.output_size, list(h.hidden_size), noisy=False,
                       layer_norm=h.layer_norm, output_layernorm=h.output_layernorm, activation=h.activation,
                       output_activation=h.output_activation)
        return NetSpec("q", self._encoder_spec(), head, None, self.output_size, 1, None, None, False)

    @staticmethod
    def rescale_action(action: torch.Tensor, low: torch.Tensor, high: torch.Tensor, output_activation: str) -> torch.Tensor:
        """actors.py:141-176 for the Tanh range [-1, 1]."""
        if output_activation != "Tanh" or low.isinf().any() or high.isinf().any():
            return action
        return (low + (high - low) * ((action - (-1.0)) / 2.0)).to(low.dtype)

    def forward(self, obs) -> torch.Tensor:
        if not isinstance(obs, torch.Tensor):
            obs = torch.as_tensor(np.asarray(obs))
        obs = obs.to(self._dev, dtype=torch.float32)
        if obs.ndim == 1:
            obs = obs.unsqueeze(0)
        obs = obs.reshape(obs.shape[0], -1).contiguous()
        rows = obs.shape[0]
        lib = _lib.load()
        need = ctypes.c_size_t(0)
        desc = ctypes.byref(self.layout.desc)
        _lib.check(lib.b2rl_actor_workspace_bytes(desc, rows, ctypes.byref(need)))
        ws = torch.empty(need.value, dtype=torch.uint8, device=self._dev)
        out = torch.empty((rows, self.output_size), dtype=torch.float32, device=self._dev)
        _lib.check(lib.b2rl_actor_forward(desc, self.buffers.params.data_ptr(), obs.data_ptr(), rows, out.data_ptr(),
                                          ws.data_ptr(), ws.numel(), _lib.stream_ptr(self._dev)))
        return out

    __call__ = forward


class ContinuousQNetwork(EvolvableNetwork):
    kind = "q"

    def __init__(self, observation_space, action_space, encoder_cls=None, encoder_config: dict | None = None,
                 head_config: dict | None = None, min_latent_dim: int = 8, max_latent_dim: int = 128, latent_dim: int = 32,
                 simba: bool = False, normalize_actions: bool = False, recurrent: bool = False, device: str = "cuda",
                 random_seed: int | None = None) -> None:
        _mlp_only(observation_space)
        if normalize_actions:
            raise NotImplementedError("normalize_actions is not implemented in the CUDA critic")
        encoder_config = copy.deepcopy(encoder_config) if encoder_config is not None else \
            dict(hidden_size=[64, 64], output_activation="ReLU")
        encoder_config["layer_norm"] = False                       # q_networks.py:349-367
        super().__init__(observation_space, encoder_cls=encoder_cls, encoder_config=encoder_config,
                         action_space=action_space, min_latent_dim=min_latent_dim, max_latent_dim=max_latent_dim,
                         latent_dim=latent_dim, simba=simba, recurrent=recurrent, device=device, random_seed=random_seed)
        self.normalize_actions = normalize_actions
        self.num_actions = int(action_space.shape[0])
        head_config = dict(head_config) if head_config is not None else dict(hidden_size=[64])
        head_config["output_activation"] = None
        self.head_net = EvolvableMLP(num_inputs=self.latent_dim + self.num_actions, num_outputs=1, name="value", device=device,
                                     random_seed=random_seed, **_head_kwargs(head_config, EvolvableMLP))
        self._finish_init()

    def _net_spec(self) -> NetSpec:
        h = self.head_net
        h.num_inputs = self.latent_dim + self.num_actions             # EvolvableNetwork._rebuild resets it to latent_dim
        head = MlpSpec("head_net.model.", "value", self.latent_dim + self.num_actions, 1, list(h.hidden_size), noisy=False,
                       layer_norm=h.layer_norm, output_layernorm=h.output_layernorm, activation=h.activation,
                       output_activation=h.output_activation)
        return NetSpec("q", self._encoder_spec(), head, None, 1, 1, None, None, False)


class MultiInputContinuousQNetwork:
    """``ContinuousQNetwork`` over a ``Dict`` of vector observation spaces — MADDPG's centralised critic
    (maddpg.py:329-337): ``EvolvableMultiInput`` with no feature nets reduces to ``final_dense`` Linear(sum obs -> latent)
    + ReLU over the raw vectors concatenated in key order (modules/multi_input.py:404-465), then
    ``cat(latent, actions)`` -> LayerNorm MLP head -> 1 (q_networks.py:424-425).  State-dict keys are the reference's
    (``encoder.final_dense.weight``, ``head_net.model.value_linear_layer_1.weight`` ...).  A fixed architecture over one
    flat HBM parameter buffer: forward / backward run inside ``b2rl_maddpg_learn``; architecture mutations of
    multi-agent networks are not implemented on the CUDA path."""

    def __init__(self, observation_space, action_space, latent_dim: int = 32, head_config: dict | None = None,
                 device: str = "cuda", random_seed: int | None = None) -> None:
        from ..engine import NetBuffers
        from .init import init_state_dict
        from .spec import DenseSpec, FlatLayout
        if not isinstance(observation_space, spaces.Dict):
            raise TypeError("observation_space must be a Dict of the agents' observation spaces")
        for k, sp in observation_space.items():
            if not (isinstance(sp, spaces.Box) and len(sp.shape) == 1):
                raise NotImplementedError(f"sub-space {k!r}: only 1-D Box observations are implemented on the CUDA path")
        if not (isinstance(action_space, spaces.Box) and len(action_space.shape) == 1):
            raise NotImplementedError("the CUDA critic takes the concatenated continuous actions (1-D Box)")
        self.observation_space, self.action_space = observation_space, action_space
        self.device, self._dev = device, _lib.as_device(device)
        self.latent_dim = int(latent_dim)
        self.num_inputs = int(sum(sp.shape[0] for sp in observation_space.values()))
        self.num_actions = int(action_space.shape[0])
        head_config = dict(head_config) if head_config is not None else dict(hidden_size=[64])
        self.head_config = head_config
        self.hidden_size = list(head_config.get("hidden_size", [64]))
        self.activation = head_config.get("activation", "ReLU")
        self.layer_norm = bool(head_config.get("layer_norm", True))
        self.random_seed = random_seed
        head = MlpSpec("head_net.model.", "value", self.latent_dim + self.num_actions, 1, self.hidden_size, noisy=False,
                       layer_norm=self.layer_norm, output_layernorm=False, activation=self.activation, output_activation=None)
        spec = NetSpec("q", DenseSpec("encoder.final_dense", self.num_inputs, self.latent_dim, "ReLU"), head, None, 1, 1,
                       None, None, False)
        self.layout = FlatLayout(spec)
        self.buffers = NetBuffers(self.layout, self._dev)
        self.buffers.load_state_dict(init_state_dict(self.layout, output_vanish_heads=True), strict=False)

    @property
    def init_dict(self) -> dict:
        return dict(observation_space=self.observation_space, action_space=self.action_space, latent_dim=self.latent_dim,
                    head_config=copy.deepcopy(self.head_config), device=self.device, random_seed=self.random_seed)

    def state_dict(self):
        return self.buffers.state_dict()

    def load_state_dict(self, sd, strict: bool = True) -> None:
        self.buffers.load_state_dict(sd, strict=strict)

    def clone(self):
        c = type(self)(**self.init_dict)
        c.buffers.copy_from(self.buffers)
        return c
