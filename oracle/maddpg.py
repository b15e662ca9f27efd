"""ORACLE (test infrastructure, never imported by the product path) — SURVEY §8(f) rank 4.

CPU restatements of
  * ``MultiAgentReplayBuffer``       agilerl/components/multi_agent_replay_buffer.py:30-242
    (deque(maxlen) of per-step experiences; ``sample`` = ``random.sample(memory, k)`` from Python's GLOBAL ``random``
    stream, then per field / per agent stacking, binary fields through ``uint8`` when the batch holds no NaN, every
    leaf ``.float()`` on the way out — utils/algo_utils.py:743-770)
  * ``MADDPG.learn / _learn_individual / soft_update``    agilerl/algorithms/maddpg.py:571-740
in plain functional torch-fp32, driven from reference ``state_dict``s.  Pinned bit-exactly against the unmodified
reference executed through ``oracle/refshim`` by ``tests/golden/make_golden.py`` (fixtures ``maddpg_vector.npz``,
``ma_replay.npz``).

Networks as ``MADDPG.__init__`` builds them for vector observations (maddpg.py:272-350; verified on the real objects):
  * actor_i  = ``DeterministicActor``: MLP encoder WITH LayerNorm (affine on hidden layers, plain on the output) ->
               LayerNorm MLP head, Tanh                                       networks/actors.py:78-210
  * critic_i = ``ContinuousQNetwork`` over the Dict of ALL agents' observation spaces: ``EvolvableMultiInput`` with no
               feature nets = ``final_dense`` Linear(sum obs -> latent) + ReLU over the concatenated raw vectors
               (modules/multi_input.py:404-465), then cat(latent, ALL agents' actions) -> LayerNorm MLP head -> 1
               (networks/q_networks.py:424-425)
Quirks kept literally: next actions of every agent come from the target actors BEFORE any update of this call; agent
i's actor step goes through its UPDATED critic; NaN rewards -> 0, NaN dones -> 1 (then ``uint8``); all soft updates
run after every agent has stepped; ``nn.MSELoss()`` mean; Adam defaults, no clipping.
"""
from __future__ import annotations

import random

import numpy as np
import torch
import torch.nn.functional as F

from .nets import MlpSpec, mlp_forward

BINARY_FIELDS = ("done", "termination", "terminated", "truncation", "truncated")


# ------------------------------------------------------------------------------------------------------------
# replay
class OracleMAReplay:
    """Ring restatement of the reference's ``deque(maxlen=memory_size)``: logical element j of the deque lives in slot
    ``(head + j) % memory_size``; ``random.sample(deque, k)`` picks by position, so ``random.sample(range(n), k)``
    consumes the global stream identically and yields those positions."""

    def __init__(self, memory_size: int, field_names, agent_ids):
        self.memory_size, self.field_names, self.agent_ids = memory_size, list(field_names), list(agent_ids)
        self.slots: list = [None] * memory_size
        self.head = 0            # slot of the oldest element
        self.n = 0
        self.counter = 0

    def __len__(self):
        return self.n

    def _add(self, *args):
        slot = (self.head + self.n) % self.memory_size
        if self.n == self.memory_size:          # deque(maxlen): the oldest element falls out on the left
            slot = self.head
            self.head = (self.head + 1) % self.memory_size
        else:
            self.n += 1
        self.slots[slot] = args
        self.counter += 1

    def save_to_memory(self, *args, is_vectorised: bool = False):
        if not is_vectorised:                   # multi_agent_replay_buffer.py:171-179
            self._add(*args)
            return
        num = len(next(iter(args[0].values())))  # :196 — number of vectorised environments
        for i in range(num):                    # :197-211, :222-224
            self._add(*[{k: np.asarray(v[i]) for k, v in arg.items()} for arg in args])

    def sample_positions(self, batch_size: int):
        return random.sample(range(self.n), k=batch_size)       # :166

    def sample(self, batch_size: int):
        pos = self.sample_positions(batch_size)
        exps = [self.slots[(self.head + p) % self.memory_size] for p in pos]
        out = []
        for fi, field in enumerate(self.field_names):
            d = {}
            for aid in self.agent_ids:
                ts = np.array([np.asarray(e[fi][aid]) for e in exps])     # stack_transitions :82-84
                if ts.ndim == 1:
                    ts = np.expand_dims(ts, axis=1)
                if field in BINARY_FIELDS and not np.isnan(ts).any():     # :147-148
                    ts = ts.astype(np.uint8)
                d[aid] = torch.as_tensor(ts).float()                      # obs_to_tensor
            out.append(d)
        return tuple(out)


# ------------------------------------------------------------------------------------------------------------
# networks
def actor_specs(obs_dim: int, act_dim: int, latent_dim: int = 32, encoder_hidden=(64, 64), head_hidden=(64,),
                activation: str = "ReLU") -> tuple[MlpSpec, MlpSpec]:
    enc = MlpSpec("encoder.model.", "encoder", obs_dim, latent_dim, list(encoder_hidden), noisy=False, layer_norm=True,
                  output_layernorm=True, activation=activation, output_activation=activation)
    head = MlpSpec("head_net.model.", "actor", latent_dim, act_dim, list(head_hidden), noisy=False, layer_norm=True,
                   activation=activation, output_activation="Tanh")
    return enc, head


def critic_head_spec(total_act: int, latent_dim: int = 32, head_hidden=(64,), activation: str = "ReLU") -> MlpSpec:
    return MlpSpec("head_net.model.", "value", latent_dim + total_act, 1, list(head_hidden), noisy=False, layer_norm=True,
                   activation=activation, output_activation=None)


def actor_forward(sd, specs, obs: torch.Tensor) -> torch.Tensor:
    enc, head = specs
    return mlp_forward(sd, head, mlp_forward(sd, enc, obs.float()))


def critic_forward(sd, head: MlpSpec, obs_list, stacked_actions: torch.Tensor) -> torch.Tensor:
    """multi_input.py:404-465 with only vector sub-spaces (features = cat of the raw observations in agent order ->
    final_dense -> ReLU), then q_networks.py:424-425."""
    feats = torch.cat([o.float() for o in obs_list], dim=1)
    latent = F.relu(F.linear(feats, sd["encoder.final_dense.weight"], sd["encoder.final_dense.bias"]))
    return mlp_forward(sd, head, torch.cat([latent, stacked_actions], dim=-1))


def _leaf(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


class OracleMADDPG:
    def __init__(self, agent_ids, a_specs: dict, c_head: MlpSpec, actor_sds: dict, actor_target_sds: dict, critic_sds: dict,
                 critic_target_sds: dict, *, gamma=0.95, tau=0.01, lr_actor=1e-3, lr_critic=1e-2):
        self.agent_ids, self.a_specs, self.c_head = list(agent_ids), a_specs, c_head
        self.actors = {a: _leaf(actor_sds[a]) for a in agent_ids}
        self.critics = {a: _leaf(critic_sds[a]) for a in agent_ids}
        self.actor_targets = {a: {k: v.clone() for k, v in actor_target_sds[a].items()} for a in agent_ids}
        self.critic_targets = {a: {k: v.clone() for k, v in critic_target_sds[a].items()} for a in agent_ids}
        self.gamma, self.tau = gamma, tau
        self.opt_actor = {a: torch.optim.Adam(list(self.actors[a].values()), lr=lr_actor) for a in agent_ids}
        self.opt_critic = {a: torch.optim.Adam(list(self.critics[a].values()), lr=lr_critic) for a in agent_ids}
        self.last_grads: dict = {}

    def _soft(self, net, target):
        with torch.no_grad():
            for k in net:
                target[k].copy_(self.tau * net[k].data + (1.0 - self.tau) * target[k])

    def learn(self, experiences):
        """maddpg.py:571-628.  ``experiences`` = (states, actions, rewards, next_states, dones), dicts by agent id."""
        states, actions, rewards, next_states, dones = experiences
        rewards, dones = dict(rewards), dict(dones)
        ids = self.agent_ids
        with torch.no_grad():
            next_actions = [actor_forward(self.actor_targets[a], self.a_specs[a], next_states[a]) for a in ids]
        stacked_actions = torch.cat([actions[a] for a in ids], dim=1)
        stacked_next_actions = torch.cat(next_actions, dim=1)
        obs_list, next_obs_list = [states[a] for a in ids], [next_states[a] for a in ids]
        out = {}
        for a in ids:                                                     # _learn_individual :630-731
            q = critic_forward(self.critics[a], self.c_head, obs_list, stacked_actions)
            with torch.no_grad():
                q_next = critic_forward(self.critic_targets[a], self.c_head, next_obs_list, stacked_next_actions)
            r = torch.where(torch.isnan(rewards[a]), torch.full_like(rewards[a], 0), rewards[a]).to(torch.float32)
            d = torch.where(torch.isnan(dones[a]), torch.full_like(dones[a], 1), dones[a]).to(torch.uint8)
            rewards[a], dones[a] = r, d
            y = r + (1 - d) * self.gamma * q_next
            critic_loss = F.mse_loss(q, y)
            self.opt_critic[a].zero_grad()
            critic_loss.backward()
            self.last_grads.update({f"critic/{a}/{k}": v.grad.detach().clone() for k, v in self.critics[a].items()})
            self.opt_critic[a].step()
            action = actor_forward(self.actors[a], self.a_specs[a], states[a])
            detached = dict(actions)
            detached[a] = action
            stacked_detached = torch.cat([detached[b] for b in ids], dim=1)
            actor_loss = -critic_forward(self.critics[a], self.c_head, obs_list, stacked_detached).mean()
            self.opt_actor[a].zero_grad()
            actor_loss.backward()
            self.last_grads.update({f"actor/{a}/{k}": v.grad.detach().clone() for k, v in self.actors[a].items()})
            self.opt_actor[a].step()
            out[a] = (actor_loss.item(), critic_loss.item())
        for a in ids:
            self._soft(self.actors[a], self.actor_targets[a])
            self._soft(self.critics[a], self.critic_targets[a])
        return out
