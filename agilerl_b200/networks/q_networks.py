"""``QNetwork`` / ``RainbowQNetwork`` — mirrors of agilerl/networks/q_networks.py:18-299.

``forward(obs)`` returns expected Q-values computed by the CUDA forward (encoder GEMMs, fused
head).  ``RainbowQNetwork.forward(obs, q=False)`` (the per-atom distribution used only inside the
reference's ``_dqn_loss``) is not exposed: the loss path is a single fused call
(``b2rl_rainbow_loss``)."""
from __future__ import annotations

import numpy as np
import torch

from ..engine import LearnEngine
from ..compat import spaces
from ..modules.mlp import EvolvableMLP
from .base import EvolvableNetwork
from .custom_modules import DuelingDistributionalMLP
from .spec import MlpSpec, NetSpec


def _head_kwargs(head_config: dict, allowed_cls) -> dict:
    import inspect
    ok = set(inspect.signature(allowed_cls.__init__).parameters)
    return {k: v for k, v in head_config.items() if k in ok and k not in ("num_inputs", "num_outputs", "device", "name",
                                                                           "num_atoms", "support", "random_seed")}


class _ForwardMixin:
    def _engine(self) -> LearnEngine:
        if getattr(self, "_engine_cache", None) is None:
            self._engine_cache = LearnEngine(self.layout, self.buffers, self.buffers)
        return self._engine_cache

    def _prep_obs(self, obs) -> torch.Tensor:
        if not isinstance(obs, torch.Tensor):
            obs = torch.as_tensor(np.asarray(obs))
        obs = obs.to(self._dev)
        shape = tuple(self.observation_space.shape)
        if obs.ndim == len(shape):
            obs = obs.unsqueeze(0)
        elif obs.ndim == len(shape) + 2:
            obs = obs.reshape(-1, *shape)
        return obs.reshape(obs.shape[0], -1) if len(shape) != 3 else obs


class QNetwork(_ForwardMixin, EvolvableNetwork):
    kind = "q"

    def __init__(self, observation_space, action_space, encoder_cls=None, encoder_config: dict | None = None,
                 head_config: dict | None = None, min_latent_dim: int = 8, max_latent_dim: int = 128,
                 latent_dim: int = 32, simba: bool = False, recurrent: bool = False, device: str = "cuda",
                 random_seed: int | None = None) -> None:
        super().__init__(observation_space, encoder_cls=encoder_cls, encoder_config=encoder_config,
                         action_space=action_space, min_latent_dim=min_latent_dim, max_latent_dim=max_latent_dim,
                         latent_dim=latent_dim, simba=simba, recurrent=recurrent, device=device, random_seed=random_seed)
        if not isinstance(action_space, (spaces.Discrete, spaces.MultiDiscrete)):
            raise ValueError("Action space must be either Discrete or MultiDiscrete")
        head_config = dict(head_config) if head_config is not None else dict(hidden_size=[32])
        head_config["output_activation"] = None
        self.num_actions = int(spaces.flatdim(action_space))
        self.head_net = EvolvableMLP(num_inputs=self.latent_dim, num_outputs=self.num_actions, name="value",
                                     device=device, random_seed=random_seed, **_head_kwargs(head_config, EvolvableMLP))
        self._finish_init()

    def _net_spec(self) -> NetSpec:
        h = self.head_net
        low, high, u8 = self._obs_normalisation()
        val = MlpSpec("head_net.model.", "value", self.latent_dim, self.num_actions, list(h.hidden_size), noisy=h.noisy,
                      layer_norm=h.layer_norm, output_layernorm=h.output_layernorm, activation=h.activation,
                      output_activation=h.output_activation)
        return NetSpec("q", self._encoder_spec(), val, None, self.num_actions, 1, low, high, u8)

    def forward(self, obs) -> torch.Tensor:
        return self._engine().q_values(self.buffers, self._prep_obs(obs), None, use_noise=False)

    __call__ = forward


class RainbowQNetwork(_ForwardMixin, EvolvableNetwork):
    kind = "rainbow"

    def __init__(self, observation_space, action_space, support: torch.Tensor, num_atoms: int = 51,
                 noise_std: float = 0.5, encoder_config: dict | None = None, head_config: dict | None = None,
                 min_latent_dim: int = 8, max_latent_dim: int = 128, latent_dim: int = 32, device: str = "cuda",
                 random_seed: int | None = None) -> None:
        if isinstance(observation_space, spaces.Box) and len(observation_space.shape) != 3:
            # q_networks.py:189-206 (the MLP encoder stays a plain Linear stack: `noisy` is never set)
            encoder_config = dict(encoder_config) if encoder_config is not None else \
                dict(hidden_size=[64, 64], output_activation="ReLU", layer_norm=True, output_vanish=False)
            encoder_config["output_activation"] = encoder_config.get("activation", "ReLU")
            encoder_config["output_vanish"] = False
            encoder_config["init_layers"] = False
            encoder_config["layer_norm"] = True
        super().__init__(observation_space, encoder_config=encoder_config, action_space=action_space,
                         min_latent_dim=min_latent_dim, max_latent_dim=max_latent_dim, latent_dim=latent_dim,
                         device=device, random_seed=random_seed)
        if not isinstance(action_space, (spaces.Discrete, spaces.MultiDiscrete)):
            raise ValueError("Action space must be either Discrete or MultiDiscrete")
        head_config = dict(head_config) if head_config is not None else dict(hidden_size=[16])
        head_config["output_activation"] = None
        for arg in ("noisy", "init_layers", "layer_norm", "output_vanish"):
            head_config.pop(arg, None)
        head_config["noise_std"] = noise_std
        self.num_actions = int(spaces.flatdim(action_space))
        self.num_atoms, self.noise_std = num_atoms, noise_std
        self.support = support.to(self._dev, dtype=torch.float32).contiguous()
        self.head_net = DuelingDistributionalMLP(num_inputs=self.latent_dim, num_outputs=self.num_actions,
                                                 num_atoms=num_atoms, support=self.support, device=device,
                                                 random_seed=random_seed,
                                                 **_head_kwargs(head_config, DuelingDistributionalMLP))
        self._finish_init()
        self.reset_noise()

    def _net_spec(self) -> NetSpec:
        h = self.head_net
        low, high, u8 = self._obs_normalisation()
        val = MlpSpec("head_net.model.", "value", self.latent_dim, self.num_atoms, list(h.hidden_size), noisy=h.noisy,
                      layer_norm=h.layer_norm, output_layernorm=h.output_layernorm, activation=h.activation,
                      output_activation=h.output_activation)
        adv = MlpSpec("head_net.advantage_net.", "advantage", self.latent_dim, self.num_actions * self.num_atoms,
                      list(h.hidden_size), noisy=h.noisy, layer_norm=h.layer_norm, output_layernorm=h.output_layernorm,
                      activation=h.activation, output_activation=h.output_activation)
        return NetSpec("rainbow", self._encoder_spec(), val, adv, self.num_actions, self.num_atoms, low, high, u8)

    def forward(self, obs, q: bool = True, log: bool = False) -> torch.Tensor:
        """q_networks.py:265-284: expected Q-values [rows, A] (q=True), per-atom distributions [rows, A, N]
        (q=False: softmax + clamp 1e-3) or log-probabilities (log=True, which takes precedence like
        custom_modules.py:154-156)."""
        if log or not q:
            return self._engine().distributions(self.buffers, self._prep_obs(obs), use_noise=self.training, log=log)
        return self._engine().q_values(self.buffers, self._prep_obs(obs), self.support, use_noise=self.training)

    __call__ = forward

    def reset_noise(self, normals: torch.Tensor | None = None) -> None:
        """NoisyLinear.reset_noise for every noisy layer, traversal order (custom_components.py:116-131)."""
        self._engine().reset_noise(self.buffers, normals)

    def recreate_network(self) -> None:
        super().recreate_network()
        self.reset_noise()
