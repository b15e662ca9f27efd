"""``DQN`` — drop-in for agilerl/algorithms/dqn.py:30-409: (double) Q-learning with MSE loss, Adam
(no gradient clipping, quirk Q14) and Polyak target update; ``learn(experiences) -> float``
(:326-347), ``get_action(obs, epsilon, action_mask)`` (:193-272)."""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from ..engine import LearnEngine
from ..networks.q_networks import QNetwork
from .core.base import RLAlgorithm
from .core.optimizer_wrapper import OptimizerWrapper
from .core.registry import HyperparameterConfig, NetworkGroup, OptimizerConfig
from .dqn_rainbow import RainbowDQN, obs_channels_to_first


class DQN(RLAlgorithm):
    def __init__(self, observation_space, action_space, index: int = 0, hp_config: HyperparameterConfig | None = None,
                 net_config: dict | None = None, batch_size: int = 64, lr: float = 1e-4, learn_step: int = 5,
                 gamma: float = 0.99, tau: float = 1e-3, mut: str | None = None, double: bool = False,
                 normalize_images: bool = True, actor_network=None, device: str = "cuda", accelerator: Any | None = None,
                 cudagraphs: bool = False, wrap: bool = True) -> None:
        super().__init__(observation_space, action_space, index=index, hp_config=hp_config, device=device,
                         accelerator=accelerator, normalize_images=normalize_images, name="DQN")
        assert learn_step >= 1, "Learn step must be greater than or equal to one."
        assert isinstance(learn_step, int), "Learn step rate must be an integer."
        assert isinstance(batch_size, int), "Batch size must be an integer."
        assert batch_size >= 1, "Batch size must be greater than or equal to one."
        assert isinstance(lr, float), "Learning rate must be a float."
        assert lr > 0, "Learning rate must be greater than zero."
        assert isinstance(gamma, (float, int, torch.Tensor)), "Gamma must be a float."
        assert isinstance(tau, float), "Tau must be a float."
        assert tau > 0, "Tau must be greater than zero."
        assert isinstance(double, bool), "Double Q-learning flag must be boolean value True or False."
        self.batch_size, self.lr, self.learn_step, self.gamma, self.tau = batch_size, lr, learn_step, gamma, tau
        self.mut, self.double, self.net_config, self.cudagraphs = mut, double, net_config, False
        if actor_network is not None:
            if not isinstance(actor_network, QNetwork):
                raise TypeError(f"'actor_network' argument is of type {type(actor_network)}, but must be of type "
                                "EvolvableModule.")
            self.actor, self.actor_target = actor_network.clone(), actor_network.clone()
        else:
            net_config = {} if net_config is None else dict(net_config)
            self.actor = QNetwork(observation_space, action_space, device=self.device, **net_config)
            self.actor_target = QNetwork(observation_space, action_space, device=self.device, **net_config)
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.register_network_group(NetworkGroup(eval_network="actor", shared_networks="actor_target", policy=True))
        self.registry.register_optimizer(OptimizerConfig(name="optimizer", networks=["actor"], lr="lr"))
        self._bind_engine()

    _bind_engine = RainbowDQN._bind_engine
    reinit_optimizers = RainbowDQN.reinit_optimizers
    _after_network_swap = RainbowDQN._after_network_swap
    _copy_networks_to = RainbowDQN._copy_networks_to
    __setattr__ = RainbowDQN.__setattr__
    soft_update = RainbowDQN.soft_update

    def get_action(self, obs, epsilon: float = 0.0, action_mask: np.ndarray | None = None) -> np.ndarray:
        """dqn.py:193-272: masked epsilon-greedy (random draws from torch's global generator)."""
        obs = self.preprocess_observation(obs)
        q_values = self.actor(obs)
        if action_mask is not None:
            action_mask = (np.stack(action_mask) if getattr(action_mask, "dtype", None) == object or
                           isinstance(action_mask, list) else action_mask)
            mask = torch.as_tensor(action_mask, device=q_values.device, dtype=q_values.dtype)
        else:
            mask = torch.ones_like(q_values)
        random_actions = torch.argmax(torch.rand_like(q_values) * mask, dim=-1)
        policy_actions = torch.argmax(q_values.masked_fill((1 - mask).bool(), float("-inf")), dim=-1)
        use_policy = torch.empty(policy_actions.shape, device=q_values.device).uniform_().gt(epsilon)
        return torch.where(use_policy, policy_actions, random_actions).cpu().numpy()

    def learn(self, experiences) -> float:
        """dqn.py:326-347 (update :274-324 + soft_update) as one fused launch sequence."""
        # the reference's update works on whatever rows the batch holds (dqn.py:274-324 has no ``range(batch_size)``
        # indexing — that is Rainbow's quirk Q17): the batch, not ``self.batch_size``, sets the row count
        r = experiences["reward"]
        rows = int(r.numel()) if isinstance(r, torch.Tensor) else int(np.asarray(r).size)
        loss = self.engine.dqn_learn(experiences, B=rows,
                                     hp=dict(gamma=self.gamma, lr=self.lr, tau=self.tau), double=self.double)
        return loss.item()

    def test(self, env, swap_channels: bool = False, max_steps: int | None = None, loop: int = 3) -> float:
        """dqn.py:360-409."""
        self.set_training_mode(False)
        rewards = []
        num_envs = env.num_envs if hasattr(env, "num_envs") else 1
        for _ in range(loop):
            obs, info = env.reset()
            scores, completed, finished = np.zeros(num_envs), np.zeros(num_envs), np.zeros(num_envs)
            step = 0
            while not np.all(finished):
                if swap_channels:
                    obs = obs_channels_to_first(obs)
                action = self.get_action(obs, epsilon=0.0, action_mask=info.get("action_mask", None))
                obs, reward, done, trunc, info = env.step(action)
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
