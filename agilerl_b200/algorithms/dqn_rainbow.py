"""``RainbowDQN`` — drop-in for agilerl/algorithms/dqn_rainbow.py:30-555.

Same constructor (:77-103) and asserts (:115-147), same attributes, ``learn(experiences,
n_experiences=None, per=False) -> (loss: float, idxs, priorities: np.ndarray | None)``
(:369-490), ``get_action`` (:239-282), ``test`` (:503-555), ``soft_update`` (:492-501).  The body
of ``learn`` — three forwards, C51 projection, cross-entropy, backward, ``clip_grad_norm_(10)``,
Adam, Polyak update, noise reset — runs in libb2rl.so (csrc/nn.cu).

Shape quirks of the reference are reproduced rather than fixed: ``weights`` of
shape [B,1] give ``mean(l) * mean(w)``; n-step ``reward``/``done`` of shape [B,1,1] (what the
unchanged driver produces) give the summed-over-batch projection.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .. import _lib
from ..compat import spaces
from ..engine import LearnEngine
from ..networks.q_networks import RainbowQNetwork
from .core.base import RLAlgorithm
from .core.optimizer_wrapper import OptimizerWrapper
from .core.registry import HyperparameterConfig, NetworkGroup, OptimizerConfig


def obs_channels_to_first(obs: np.ndarray) -> np.ndarray:
    """utils/algo_utils.py: [H, W, C] -> [C, H, W] (batched or not)."""
    return np.moveaxis(obs, -1, -3)


class RainbowDQN(RLAlgorithm):
    def __init__(self, observation_space, action_space, index: int = 0, hp_config: HyperparameterConfig | None = None,
                 net_config: dict | None = None, batch_size: int = 64, lr: float = 1e-4, learn_step: int = 5,
                 gamma: float = 0.99, tau: float = 1e-3, beta: float = 0.4, prior_eps: float = 1e-6,
                 num_atoms: int = 51, v_min: float = 0, v_max: float = 200, noise_std: float = 0.5, n_step: int = 3,
                 mut: str | None = None, normalize_images: bool = True, combined_reward: bool = False,
                 actor_network=None, device: str = "cuda", accelerator: Any | None = None, wrap: bool = True) -> None:
        super().__init__(observation_space, action_space, index=index, hp_config=hp_config, device=device,
                         accelerator=accelerator, normalize_images=normalize_images, name="Rainbow DQN")
        assert learn_step >= 1, "Learn step must be greater than or equal to one."
        assert isinstance(learn_step, int), "Learn step rate must be an integer."
        assert isinstance(batch_size, int), "Batch size must be an integer."
        assert batch_size >= 1, "Batch size must be greater than or equal to one."
        assert isinstance(lr, float), "Learning rate must be a float."
        assert lr > 0, "Learning rate must be greater than zero."
        assert isinstance(gamma, (float, int, torch.Tensor)), "Gamma must be a float."
        assert isinstance(tau, float), "Tau must be a float."
        assert tau > 0, "Tau must be greater than zero."
        assert isinstance(prior_eps, float), "Minimum priority for sampling must be a float."
        assert prior_eps > 0, "Minimum priority for sampling must be greater than zero."
        assert isinstance(num_atoms, int), "Number of atoms must be an integer."
        assert num_atoms >= 1, "Number of atoms must be greater than or equal to one."
        assert isinstance(v_min, (float, int)), "Minimum value of support must be a float."
        assert isinstance(v_max, (float, int)), "Maximum value of support must be a float."
        assert v_max >= v_min, "Maximum value of support must be greater than or equal to minimum value."
        assert isinstance(n_step, int), "Step number must be an integer."
        assert n_step >= 1, "Step number must be greater than or equal to one."
        assert isinstance(wrap, bool), "Wrap models flag must be boolean value True or False."
        if not normalize_images and len(observation_space.shape) == 3:
            raise NotImplementedError("normalize_images=False for image observations is not wired to the CUDA loader")

        self.batch_size, self.learn_step, self.lr = batch_size, learn_step, lr
        self.gamma, self.tau, self.beta, self.prior_eps = gamma, tau, beta, prior_eps
        self.num_atoms, self.net_config = num_atoms, net_config
        self.v_min, self.v_max, self.n_step, self.mut = v_min, v_max, n_step, mut
        self.combined_reward, self.noise_std = combined_reward, noise_std
        # dqn_rainbow.py:165-171 — linspace evaluated on the host (fp32) and copied, so the atoms are
        # bit-identical to the reference's CPU tensor
        self.support = torch.linspace(self.v_min, self.v_max, self.num_atoms).to(self._dev)
        self.delta_z = (self.v_max - self.v_min) / (self.num_atoms - 1)

        if actor_network is not None:
            if not isinstance(actor_network, RainbowQNetwork):
                raise TypeError(f"'actor_network' argument is of type {type(actor_network)}, but must be of type "
                                "EvolvableModule.")
            self.actor, self.actor_target = actor_network.clone(), actor_network.clone()
        else:
            net_config = {} if net_config is None else dict(net_config)
            head_config = dict(net_config.get("head_config", None) or {})
            head_config = dict(hidden_size=head_config.get("hidden_size", [64]), noise_std=self.noise_std,
                               output_activation="ReLU", min_mlp_nodes=head_config.get("min_mlp_nodes", 16),
                               max_mlp_nodes=head_config.get("max_mlp_nodes", 500),
                               **{k: v for k, v in head_config.items() if k in ("activation", "min_hidden_layers",
                                                                                  "max_hidden_layers")})
            net_config["head_config"] = head_config
            self.net_config = net_config

            def create_actor():
                return RainbowQNetwork(observation_space=observation_space, action_space=action_space,
                                       support=self.support, num_atoms=self.num_atoms, noise_std=self.noise_std,
                                       device=self.device, **net_config)
            self.actor = create_actor()
            self.actor_target = create_actor()
        self.actor_target.load_state_dict(self.actor.state_dict())       # :218
        self.actor.train(); self.actor_target.train()
        self.register_network_group(NetworkGroup(eval_network="actor", shared_networks="actor_target", policy=True))
        self.registry.register_optimizer(OptimizerConfig(name="optimizer", networks=["actor"], lr="lr"))
        self._bind_engine()

    # -- engine plumbing ----------------------------------------------------------------------------
    def _bind_engine(self, keep_state: dict | None = None) -> None:
        self.engine = LearnEngine(self.actor.layout, self.actor.buffers, self.actor_target.buffers)
        self.engine.philox_seed = 0xB200 + 7919 * int(self.index)
        self.optimizer = OptimizerWrapper(torch.optim.Adam, networks=self.actor, lr=self.lr, engine=self.engine)
        if keep_state is not None:
            self.optimizer.load_state_dict(keep_state)

    def reinit_optimizers(self, optimizer=None) -> None:
        """Fresh Adam state (mutation.py:441-450 / reinit after architecture changes)."""
        self._bind_engine()

    def _after_network_swap(self) -> None:
        self._bind_engine()

    def _after_hyperparameter_restore(self) -> None:
        self.support = torch.linspace(self.v_min, self.v_max, self.num_atoms).to(self._dev)
        self.delta_z = (self.v_max - self.v_min) / (self.num_atoms - 1)

    def _copy_networks_to(self, clone) -> None:
        clone.actor, clone.actor_target = self.actor.clone(), self.actor_target.clone()
        clone._bind_engine(keep_state=self.optimizer.state_dict())

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name == "lr" and "optimizer" in self.__dict__:
            self.optimizer.lr = value

    # -- acting -------------------------------------------------------------------------------------
    def get_action(self, obs, action_mask: np.ndarray | None = None, training: bool = True, *args, **kwargs) -> np.ndarray:
        """dqn_rainbow.py:239-282."""
        obs = self.preprocess_observation(obs)
        self.actor.train(mode=training)
        action_values = self.actor(obs).cpu().numpy()
        if action_mask is None:
            action = np.argmax(action_values, axis=-1)
        else:
            action_mask = (np.stack(action_mask) if getattr(action_mask, "dtype", None) == object or
                           isinstance(action_mask, list) else action_mask)
            masked = np.ma.array(action_values, mask=1 - action_mask)
            action = np.argmax(masked, axis=-1)
        self.actor.train()
        return action

    # -- learning -----------------------------------------------------------------------------------
    def _hp(self) -> dict:
        return dict(v_min=self.v_min, v_max=self.v_max, delta_z=self.delta_z, lr=self.lr, tau=self.tau,
                    prior_eps=self.prior_eps)

    @staticmethod
    def _is_driver_shape(t: torch.Tensor) -> bool:
        return t.ndim == 3          # [B,1,1] fields from storage[idxs [B,1]] (quirk Q2)

    def learn(self, experiences, n_experiences=None, per: bool = False, noise_normals=None):
        """dqn_rainbow.py:369-490."""
        n_step = n_experiences is not None
        B = self.batch_size
        passes = []
        if self.combined_reward or not n_step:
            passes.append((experiences, self.gamma, self._is_driver_shape(experiences["reward"])))
        if n_step:
            passes.append((n_experiences, self.gamma ** self.n_step, self._is_driver_shape(n_experiences["reward"])))
        weights, weights_mode, idxs = None, 0, None
        if per:
            weights = experiences["weights"]
            idxs = experiences["idxs"]
            weights_mode = 2 if weights.ndim == 2 else 1          # quirk Q1
        elif n_step:
            idxs = experiences["idxs"]
        self.engine.rainbow_learn(passes, B=B, support=self.support, weights=weights, weights_mode=weights_mode,
                                  hp=self._hp(), noise_normals=noise_normals, host_readback=True)
        # loss and priorities were copied to the host before the backward was enqueued: this waits for the
        # forward half of the step only; backward / optimiser keep running in stream order behind it
        loss, pri = self.engine.readback()
        return loss, idxs, (pri if per else None)                 # elementwise_loss + prior_eps (:487-488)

    def learn_from_buffers(self, memory, n_step_memory, overlap: bool = False, side_streams: int | None = None):
        """Fused HBM-resident gradient step (no host round trip): PER sample + learn + priority
        write-back, equivalent to train_off_policy.py:399-412 with canonical shapes.  Returns the
        loss as a DEVICE tensor.

        ``overlap=True`` leaves this agent's backward + optimiser running on its own CUDA stream
        when the call returns (the next agent's step can start meanwhile); ``synchronize()`` — or any
        later learn / get_action of this agent — joins it.  Call ``synchronize()`` before touching
        the parameters through torch (clone, state_dict, mutations) or adding to the buffers.
        ``side_streams`` (bit 0: target forward, bit 1: weight gradients on library-owned side streams)
        defaults to 3 for a sequential step and 1 for an overlapped one."""
        loss, idx, pri = self.engine.rainbow_fused_step(memory, n_step_memory, B=self.batch_size, beta=self.beta,
                                                        support=self.support, hp=self._hp(),
                                                        gamma_n=self.gamma ** self.n_step, overlap=overlap,
                                                        side_streams=side_streams)
        return loss

    def synchronize(self) -> None:
        """Make the current stream wait for an overlapped learn tail of this agent."""
        self.engine.join()

    def soft_update(self) -> None:
        """dqn_rainbow.py:492-501 (learn() already applies it inside the fused optimiser kernel)."""
        p, t = self.actor.buffers.params, self.actor_target.buffers.params
        t.copy_(self.tau * p + (1.0 - self.tau) * t)

    def test(self, env, swap_channels: bool = False, max_steps: int | None = None, loop: int = 3) -> float:
        """dqn_rainbow.py:503-555."""
        self.set_training_mode(False)
        rewards = []
        num_envs = env.num_envs if hasattr(env, "num_envs") else 1
        for _ in range(loop):
            obs, info = env.reset()
            scores = np.zeros(num_envs)
            completed = np.zeros(num_envs)
            finished = np.zeros(num_envs)
            step = 0
            while not np.all(finished):
                if swap_channels:
                    obs = obs_channels_to_first(obs)
                action = self.get_action(obs, training=False, action_mask=info.get("action_mask", None))
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
