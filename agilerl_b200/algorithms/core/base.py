"""``EvolvableAlgorithm`` / ``RLAlgorithm`` — the agent-side contract ``train_off_policy``,
``TournamentSelection`` and ``Mutations`` rely on (mirror of agilerl/algorithms/core/base.py:237-1301
for single-agent algorithms without accelerate): attributes ``index, scores, fitness, steps, mut,
registry, device, accelerator, algo``; ``clone`` (re-invoke the constructor with every attribute
that is also a constructor parameter :390-430/:855-917, then copy networks, optimiser state and
the remaining attributes); ``reinit_optimizers``; ``get_lr_names``; ``mutation_hook``;
``preprocess_observation``; checkpoints."""
from __future__ import annotations

import copy
import inspect
from typing import Any

import numpy as np
import torch

from ... import _lib
from ...compat import spaces
from .registry import HyperparameterConfig, MutationRegistry, NetworkGroup, OptimizerConfig


class EvolvableAlgorithm:
    def __init__(self, index: int, hp_config: HyperparameterConfig | None = None, device: str = "cuda",
                 accelerator: Any | None = None, torch_compiler: Any | None = None, name: str | None = None) -> None:
        assert isinstance(index, int), "Agent index must be an integer."
        if accelerator is not None:
            raise NotImplementedError(
                "accelerate/DDP wrapping is replaced by one-agent-per-GPU sharding (see DESIGN.md); pass accelerator=None")
        self.accelerator = None
        self.device = device
        self._dev = _lib.as_device(device)
        self.torch_compiler = None
        self.algo = name if name is not None else self.__class__.__name__
        self._mut = None
        self._index = index
        self.scores: list = []
        self.fitness: list = []
        self.steps: list = [0]
        self.registry = MutationRegistry(hp_config)
        # the reference's clone() deep-copies ``registry`` (copy_attributes, core/base.py:432-492); here the
        # constructor rebuilds it, so the hyper-parameter config travels as a constructor argument
        self.hp_config = self.registry.hp_config
        self.training = True

    @property
    def index(self) -> int:
        return self._index

    @index.setter
    def index(self, value: int) -> None:
        self._index = value

    @property
    def mut(self):
        return self._mut

    @mut.setter
    def mut(self, value) -> None:
        self._mut = value

    # -- registry --------------------------------------------------------------------------------
    def register_network_group(self, group: NetworkGroup) -> None:
        self.registry.register_group(group)

    def register_mutation_hook(self, hook) -> None:
        self.registry.hooks.append(hook)

    def mutation_hook(self) -> None:
        for hook in self.registry.hooks:
            hook()

    def get_lr_names(self) -> list[str]:
        return [opt.lr for opt in self.registry.optimizers]

    def reinit_optimizers(self, optimizer: OptimizerConfig | None = None) -> None:
        raise NotImplementedError

    def recompile(self) -> None:
        pass

    def set_training_mode(self, training: bool) -> None:
        self.training = training

    def clean_up(self) -> None:
        """core/base.py:1233-1240: drop the networks and optimisers (their HBM buffers go back to the allocator) — plus the
        engine state built on them (flat gradient / Adam buffers, captured graphs)."""
        for name in list(self.evolvable_attributes()):
            if name in self.__dict__:
                object.__delattr__(self, name)
        for name in [k for k in self.__dict__ if k in ("engine", "_ws", "_keep", "_plans") or k.startswith("_engine")]:
            object.__delattr__(self, name)

    # -- clone -----------------------------------------------------------------------------------
    @classmethod
    def _ctor_params(cls) -> list[str]:
        return [p for p in inspect.signature(cls.__init__).parameters if p != "self"]

    def _init_kwargs(self) -> dict:
        out = {}
        for p in self._ctor_params():
            if p in ("actor_network", "wrap"):
                continue
            if hasattr(self, p):
                out[p] = copy.deepcopy(getattr(self, p)) if p not in ("observation_space", "action_space", "device",
                                                                      "accelerator") else getattr(self, p)
        return out

    def _evolvable_attrs(self) -> list[str]:
        return self.registry.all_registered()

    def evolvable_attributes(self, networks_only: bool = False) -> dict:
        """core/base.py:790-819: the registered evolvable networks by attribute name and, unless ``networks_only``, the
        optimisers associated with them."""
        out = {name: getattr(self, name) for name in self.registry.all_registered() if hasattr(self, name)}
        if not networks_only:
            for cfg in self.registry.optimizers:
                if hasattr(self, cfg.name):
                    out[cfg.name] = getattr(self, cfg.name)
        return out

    def clone(self, index: int | None = None, wrap: bool = True):
        """core/base.py:855-917."""
        kwargs = self._init_kwargs()
        if index is not None:
            kwargs["index"] = index
        clone = type(self)(**kwargs)
        self._copy_networks_to(clone)
        skip = set(self._ctor_params()) | set(self._evolvable_attrs()) | {"registry", "optimizer", "engine", "_dev",
                                                                          "_index", "support"}
        skip |= set(getattr(self, "_clone_skip", ()))       # per-class device state rebuilt by _copy_networks_to
        for k, v in self.__dict__.items():
            if k in skip or k.startswith("_engine"):
                continue
            try:
                setattr(clone, k, copy.deepcopy(v))
            except Exception:  # noqa: BLE001 - non-copyable attribute stays as constructed
                pass
        if index is not None:
            clone.index = index
        return clone

    def _copy_networks_to(self, clone) -> None:
        raise NotImplementedError

    # -- cross-rank move (population sharding; hpo/tournament.py) -----------------------------------
    def export_state(self):
        """-> (picklable description, [device tensors]) sufficient to rebuild this agent elsewhere."""
        nets = self._evolvable_attrs()
        meta = {"init": {k: v for k, v in self._init_kwargs().items() if k not in ("device", "accelerator")},
                "nets": {n: getattr(self, n).init_dict for n in nets},
                "attrs": {"scores": list(self.scores), "fitness": list(self.fitness), "steps": list(self.steps),
                          "mut": self.mut, "index": self.index, "opt_step": self.engine.step}}
        for n in nets:      # tensors / device strings inside init dicts do not travel
            for k in ("device", "support"):
                meta["nets"][n].pop(k, None)
        tensors = []
        for n in nets:
            net = getattr(self, n)
            tensors += [net.buffers.params, net.buffers.eps]
        tensors += [self.engine.exp_avg, self.engine.exp_avg_sq]
        return meta, tensors

    @classmethod
    def from_state(cls, meta, tensors, like):
        agent = cls(device=like.device, **meta["init"])
        nets = agent._evolvable_attrs()
        for n in nets:
            cur = getattr(agent, n)
            kw = dict(meta["nets"][n]); kw["device"] = like.device
            if "support" in cur.init_dict:
                kw["support"] = agent.support
            setattr(agent, n, type(cur)(**kw))
        agent._after_network_swap()
        it = iter(tensors)
        for n in nets:
            net = getattr(agent, n)
            net.buffers.params.copy_(next(it)); net.buffers.eps.copy_(next(it))
        agent.engine.exp_avg.copy_(next(it)); agent.engine.exp_avg_sq.copy_(next(it))
        a = meta["attrs"]
        agent.scores, agent.fitness, agent.steps, agent.mut = a["scores"], a["fitness"], a["steps"], a["mut"]
        agent.index = a["index"]; agent.engine.step = a["opt_step"]
        return agent

    # -- checkpoints (core/base.py:919-1049, reduced to what this path owns) --------------------------
    def save_checkpoint(self, path: str) -> None:
        ckpt = {"init": {k: v for k, v in self._init_kwargs().items() if k not in ("accelerator",)},
                "networks": {n: {"init_dict": getattr(self, n).init_dict, "state_dict": getattr(self, n).state_dict()}
                             for n in self._evolvable_attrs()},
                "optimizer": self.optimizer.state_dict(), "scores": self.scores, "fitness": self.fitness,
                "steps": self.steps, "mut": self.mut, "index": self.index, "algo": self.algo}
        torch.save(ckpt, path)

    def load_checkpoint(self, path: str) -> None:
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        for n, d in ckpt["networks"].items():
            net = type(getattr(self, n))(**d["init_dict"])
            net.load_state_dict(d["state_dict"])
            setattr(self, n, net)
        # every saved constructor attribute comes back (the reference setattr's the whole checkpoint,
        # core/base.py:1003-1049): mutated lr / batch_size / learn_step / beta / tau / gamma / n_step / v_min / v_max ...
        for k, v in ckpt.get("init", {}).items():
            if k in ("device", "accelerator", "observation_space", "action_space", "index", "net_config"):
                continue
            if k == "hp_config":
                if v is not None:
                    self.hp_config = self.registry.hp_config = v
                continue
            setattr(self, k, v)
        self._after_hyperparameter_restore()
        self._after_network_swap()
        self.optimizer.load_state_dict(ckpt["optimizer"], strict=True)
        self.scores, self.fitness, self.steps, self.mut = ckpt["scores"], ckpt["fitness"], ckpt["steps"], ckpt["mut"]
        self.index = ckpt["index"]

    def _after_hyperparameter_restore(self) -> None:
        """Derived attributes that depend on restored scalars (e.g. the C51 support)."""

    def _after_network_swap(self) -> None:
        pass


class RLAlgorithm(EvolvableAlgorithm):
    def __init__(self, observation_space, action_space, index: int, hp_config=None, device: str = "cuda",
                 accelerator=None, torch_compiler=None, normalize_images: bool = True, name: str | None = None) -> None:
        super().__init__(index, hp_config, device, accelerator, torch_compiler, name)
        assert isinstance(observation_space, spaces.Space), "Observation space must be an instance of gymnasium.spaces.Space."
        assert isinstance(action_space, spaces.Space), "Action space must be an instance of gymnasium.spaces.Space."
        self.observation_space, self.action_space = observation_space, action_space
        self.normalize_images = normalize_images
        self.action_dim = int(spaces.flatdim(action_space)) if isinstance(action_space, spaces.Discrete) else \
            int(np.prod(action_space.shape))

    def preprocess_observation(self, observation) -> torch.Tensor:
        """core/base.py:1287-1301: to device tensor + batch-dim fix-up.  The float cast and image
        min-max normalisation (algo_utils.py:993-1022, 1131-1180) happen inside the first layer's
        loader on the GPU (uint8 LUT of exact (x-low)/(high-low))."""
        if not isinstance(observation, torch.Tensor):
            observation = torch.as_tensor(np.asarray(observation))
        obs = observation.to(self._dev, non_blocking=True)
        shape = tuple(self.observation_space.shape)
        if obs.ndim == len(shape):
            obs = obs.unsqueeze(0)
        elif obs.ndim == len(shape) + 2:
            obs = obs.reshape(-1, *shape)
        elif obs.ndim != len(shape) + 1:
            raise ValueError(f"Expected observation to have {len(shape) + 1} dimensions, got {obs.ndim}.")
        return obs
