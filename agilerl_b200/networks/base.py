"""``EvolvableNetwork`` — encoder + head with flat HBM parameter buffers.

Mirror of agilerl/networks/base.py:134-567 for the encoders the off-policy path uses
(``_build_encoder`` :505-567: ``EvolvableCNN`` for image ``Box`` spaces, ``EvolvableMLP``
otherwise).  Owns the parameter / epsilon buffers (``NetBuffers``), rebuilds and re-lays them
out after every architecture mutation while preserving the overlapping weight slices
(``EvolvableModule.preserve_parameters`` modules/base.py:471-502 and
``EvolvableCNN.shrink_preserve_parameters`` cnn.py:417-453 reduce to the same slice copy here),
and exposes ``state_dict`` with the reference's key names.
"""
from __future__ import annotations

import copy
from collections import OrderedDict
from typing import Any

import numpy as np
import torch

from .. import _lib
from ..engine import NetBuffers
from ..compat import spaces
from ..modules.base import EvolvableModule, MutationType, mutation
from ..modules.cnn import EvolvableCNN
from ..modules.mlp import EvolvableMLP
from .init import init_state_dict
from .spec import CnnSpec, FlatLayout, MlpSpec, NetSpec


def is_image_space(space) -> bool:
    """utils/evolvable_networks.py is_image_space: a 3-D Box."""
    return isinstance(space, spaces.Box) and len(space.shape) == 3


def get_default_encoder_config(observation_space) -> dict:
    """utils/evolvable_networks.py:168-217 (CNN / MLP cases)."""
    if is_image_space(observation_space):
        return dict(channel_size=[32, 32], kernel_size=[3, 3], stride_size=[1, 1], output_activation="ReLU")
    return dict(hidden_size=[64, 64], output_activation="ReLU", layer_norm=True, output_vanish=False)


class EvolvableNetwork(EvolvableModule):
    kind = "q"

    def __init__(self, observation_space, encoder_cls=None, encoder_config: dict | None = None,
                 encoder_name: str = "encoder", action_space=None, min_latent_dim: int = 8, max_latent_dim: int = 128,
                 latent_dim: int = 32, simba: bool = False, recurrent: bool = False, device: str = "cuda",
                 random_seed: int | None = None) -> None:
        super().__init__(device, random_seed)
        assert latent_dim <= max_latent_dim, "Latent dimension must be less than or equal to max latent dimension."
        assert latent_dim >= min_latent_dim, "Latent dimension must be greater than or equal to min latent dimension."
        if encoder_cls is not None or simba or recurrent:
            raise NotImplementedError("custom / SimBa / recurrent encoders are outside the CUDA hot path")
        if isinstance(observation_space, (spaces.Dict, spaces.Tuple)):
            raise NotImplementedError("Dict/Tuple observation spaces (EvolvableMultiInput) are outside the CUDA hot path")
        self._dev = _lib.as_device(device)
        self.observation_space, self.action_space = observation_space, action_space
        self.latent_dim, self.min_latent_dim, self.max_latent_dim = latent_dim, min_latent_dim, max_latent_dim
        self.encoder_cls, self.encoder_name = None, encoder_name
        self.simba, self.recurrent = simba, recurrent
        self.training = True
        encoder_config = copy.deepcopy(encoder_config) if encoder_config is not None else \
            get_default_encoder_config(observation_space)
        if encoder_config.get("output_activation") is None:      # networks/base.py:226-230
            encoder_config["output_activation"] = encoder_config.get("activation", "ReLU")
        self.encoder = self._build_encoder(encoder_config)
        self.encoder.disable_mutations(MutationType.LAYER)        # :268-270
        self.head_net: EvolvableModule | None = None
        self.buffers: NetBuffers | None = None
        self.layout: FlatLayout | None = None

    # -- construction --------------------------------------------------------------------------
    def _build_encoder(self, cfg: dict) -> EvolvableModule:
        import inspect

        def only(cls, d):
            ok = set(inspect.signature(cls.__init__).parameters)
            return {k: v for k, v in d.items() if k in ok and k not in ("input_shape", "num_inputs", "num_outputs",
                                                                         "device", "name", "random_seed")}
        if is_image_space(self.observation_space):
            for k in ("channel_size", "kernel_size", "stride_size"):
                assert k in cfg, f"Net config must contain {k}: int."
            return EvolvableCNN(input_shape=self.observation_space.shape, num_outputs=self.latent_dim,
                                device=self.device, name=self.encoder_name, random_seed=self.random_seed,
                                **only(EvolvableCNN, cfg))
        cfg = dict(cfg)
        cfg["output_layernorm"] = cfg.get("layer_norm", True)     # networks/base.py:549-552
        cfg["output_vanish"] = False
        self.flatten_obs = len(self.observation_space.shape) > 1
        return EvolvableMLP(num_inputs=int(spaces.flatdim(self.observation_space)), num_outputs=self.latent_dim,
                            device=self.device, name=self.encoder_name, random_seed=self.random_seed,
                            **only(EvolvableMLP, cfg))

    def _finish_init(self) -> None:
        """Called by subclasses once ``head_net`` exists: lay out buffers and initialise."""
        for m in (self.encoder, self.head_net):
            m.register_mutation_hook(self._on_child_mutation)
        self._mutation_hooks.append(self._on_self_mutation)
        self._rebuild(preserve=None)

    def _encoder_spec(self):
        e = self.encoder
        if isinstance(e, EvolvableCNN):
            return CnnSpec("encoder.model.", e.name, e.input_shape, list(e.channel_size), list(e.kernel_size),
                           list(e.stride_size), self.latent_dim, e.activation, e.output_activation)
        return MlpSpec("encoder.model.", e.name, e.num_inputs, self.latent_dim, list(e.hidden_size), noisy=e.noisy,
                       layer_norm=e.layer_norm, output_layernorm=e.output_layernorm, activation=e.activation,
                       output_activation=e.output_activation)

    def _net_spec(self) -> NetSpec:
        raise NotImplementedError

    def _obs_normalisation(self):
        sp = self.observation_space
        if not is_image_space(sp):
            return None, None, False
        u8 = np.dtype(sp.dtype) == np.uint8
        low, high = np.unique(sp.low), np.unique(sp.high)
        if np.isinf(sp.high).any() or np.isinf(sp.low).any():
            return None, None, u8                                  # algo_utils.py:1146-1160: bypass
        if len(low) != 1 or len(high) != 1:
            raise NotImplementedError("per-pixel low/high image bounds are not implemented in the CUDA loader")
        return float(low[0]), float(high[0]), u8

    def _rebuild(self, preserve: "OrderedDict | None") -> None:
        self.encoder.num_outputs = self.latent_dim
        self.head_net.num_inputs = self.latent_dim
        spec = self._net_spec()
        self.layout = FlatLayout(spec)
        buffers = NetBuffers(self.layout, self._dev)
        head_noise = getattr(self.head_net, "noise_std", 0.5)
        sd = init_state_dict(self.layout, noise_std=head_noise,
                             output_vanish_heads=getattr(self.head_net, "output_vanish", True),
                             init_mlp_layers=False)
        buffers.load_state_dict(sd, strict=False)
        if preserve is not None:                                    # preserve_parameters
            for k, old in preserve.items():
                if k not in self.layout.entries or self.layout.entries[k].buf != "param":
                    continue
                new = buffers.view(k)
                if old.shape == new.shape:
                    new.copy_(old)
                elif "norm" not in k:
                    sl = tuple(slice(0, min(o, n)) for o, n in zip(old.shape, new.shape))
                    new[sl] = old[sl]
        self.buffers = buffers
        self._engine_cache = None

    def _param_snapshot(self) -> "OrderedDict":
        return OrderedDict((k, self.buffers.view(k).clone()) for k in self.layout.param_keys())

    def _on_child_mutation(self) -> None:
        if self._mutation_depth == 0:
            self.recreate_network()

    def _on_self_mutation(self) -> None:
        self.recreate_network()

    def recreate_network(self) -> None:
        self._rebuild(preserve=self._param_snapshot() if self.buffers is not None else None)

    # -- mutations (networks/base.py:457-491) --------------------------------------------------------
    @mutation(MutationType.NODE)
    def add_latent_node(self, numb_new_nodes: int | None = None) -> dict[str, Any]:
        if numb_new_nodes is None:
            numb_new_nodes = int(self.rng.choice([8, 16, 32]))
        if self.latent_dim + numb_new_nodes < self.max_latent_dim:
            self.latent_dim += numb_new_nodes
        return {"numb_new_nodes": numb_new_nodes}

    @mutation(MutationType.NODE)
    def remove_latent_node(self, numb_new_nodes: int | None = None) -> dict[str, Any]:
        if numb_new_nodes is None:
            numb_new_nodes = int(self.rng.choice([8, 16, 32]))
        if self.latent_dim - numb_new_nodes > self.min_latent_dim:
            self.latent_dim -= numb_new_nodes
        return {"numb_new_nodes": numb_new_nodes}

    def get_mutation_methods(self):
        """Dotted names for nested modules ("encoder.add_channel", "head_net.add_node")."""
        out = dict(self._own_mutation_methods())
        for attr in ("encoder", "head_net"):
            mod = getattr(self, attr)
            for name, fn in mod.get_mutation_methods().items():
                out[f"{attr}.{name}"] = self._wrap_child(attr, name, fn)
        return out

    def _wrap_child(self, attr, name, fn):
        def call(*a, **k):
            out = fn(*a, **k)
            self.last_mutation_attr = f"{attr}.{getattr(self, attr).last_mutation_attr}"
            return out
        call._mutation_type = fn._mutation_type
        call.__name__ = f"{attr}.{name}"
        return call

    # -- nn.Module-like surface ----------------------------------------------------------------------
    @property
    def activation(self) -> str:
        return self.encoder.activation

    def change_activation(self, activation: str, output: bool = False) -> None:
        """networks/base.py:448-455."""
        self._mutation_depth += 1
        try:
            self.encoder.change_activation(activation, output=True)
            self.head_net.change_activation(activation, output=output)
        finally:
            self._mutation_depth -= 1
        self.recreate_network()

    @property
    def encoder_config(self) -> dict:
        return self.encoder.net_config

    @property
    def head_config(self) -> dict:
        return self.head_net.net_config

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        return self.buffers.state_dict()

    def load_state_dict(self, sd, strict: bool = True) -> None:
        self.buffers.load_state_dict(sd, strict=strict)

    def named_parameters(self):
        for k in self.layout.param_keys():
            yield k, self.buffers.view(k)

    def parameters(self):
        for _, v in self.named_parameters():
            yield v

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def to(self, device):
        if _lib.as_device(device) != self._dev:
            raise NotImplementedError("moving a network between GPUs: clone it on the target device instead")
        return self

    def clone(self):
        c = type(self)(**self.init_dict)
        c._disabled, c._disabled_names = set(self._disabled), set(self._disabled_names)
        c.rng = copy.deepcopy(self.rng)
        c.encoder.rng, c.head_net.rng = copy.deepcopy(self.encoder.rng), copy.deepcopy(self.head_net.rng)
        c.last_mutation_attr = self.last_mutation_attr
        c.buffers.copy_from(self.buffers)
        c.training = self.training
        return c
