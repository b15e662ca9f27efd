"""ORACLE (test infrastructure, never imported by the product path) — SURVEY §8(f) rank 1.

CPU restatement of the reference's DDPG / TD3 ``learn()`` in plain functional torch-fp32, driven
from reference ``state_dict``s (reference parameter names), for the round that moves these learners
onto the HBM replay.  Pinned bit-exactly against the unmodified reference executed through
``oracle/refshim`` by ``tests/golden/make_golden.py`` (fixtures ``ddpg_*.npz`` / ``td3_*.npz``).

Follows (reference file:line):
  * DeterministicActor  = MLP encoder (no LayerNorm under DDPG/TD3) -> MLP head (LayerNorm), Tanh output
                                                                                  networks/actors.py:78-210
    (``forward`` returns the head output; rescaling to the action bounds is done by get_action only)
  * ContinuousQNetwork  = MLP encoder WITHOUT LayerNorm -> cat(latent, action) -> MLP head -> 1
                                                                                  networks/q_networks.py:302-443
  * DDPG.learn                                                                    algorithms/ddpg.py:422-494
  * TD3.learn (twin critics, min target, actor on critic_1)                       algorithms/td3.py:459-545
  * quirks kept literally: the target-policy noise is drawn by ``actions.data.normal_(0, policy_noise)``
    — IN PLACE on the batch's action tensor, after the critics consumed it — clamped to +-noise_clip;
    actor and ALL target networks move only every ``policy_freq`` learn calls; Adam defaults;
    ``nn.MSELoss()`` mean; TD3's critic loss is the SUM of the two MSEs, backpropagated once.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .nets import MlpSpec, mlp_forward


def actor_specs(obs_dim: int, act_dim: int, latent_dim: int = 32, encoder_hidden=(64, 64), head_hidden=(32,),
                activation: str = "ReLU", encoder_layer_norm: bool = False) -> tuple[MlpSpec, MlpSpec]:
    """(encoder, head) of DeterministicActor as DDPG/TD3 build it (base.py:505-567, actors.py:116-147).
    DDPG/TD3 hand the actor the same encoder config as the critics, whose LayerNorm is disabled
    (ddpg.py net_config handling; verified on the real object: plain Linear->ReLU encoder)."""
    enc = MlpSpec("encoder.model.", "encoder", obs_dim, latent_dim, list(encoder_hidden), noisy=False,
                  layer_norm=encoder_layer_norm, output_layernorm=encoder_layer_norm, activation=activation,
                  output_activation=activation)
    head = MlpSpec("head_net.model.", "actor", latent_dim, act_dim, list(head_hidden), noisy=False, layer_norm=True,
                   activation=activation, output_activation="Tanh")
    return enc, head


def critic_specs(obs_dim: int, act_dim: int, latent_dim: int = 32, encoder_hidden=(64, 64), head_hidden=(64,),
                 activation: str = "ReLU") -> tuple[MlpSpec, MlpSpec]:
    """(encoder, head) of ContinuousQNetwork: the encoder drops LayerNorm (q_networks.py:349-367)."""
    enc = MlpSpec("encoder.model.", "encoder", obs_dim, latent_dim, list(encoder_hidden), noisy=False, layer_norm=False,
                  output_layernorm=False, activation=activation, output_activation=activation)
    head = MlpSpec("head_net.model.", "value", latent_dim + act_dim, 1, list(head_hidden), noisy=False, layer_norm=True,
                   activation=activation, output_activation=None)
    return enc, head


def actor_forward(sd, specs, obs: torch.Tensor) -> torch.Tensor:
    enc, head = specs
    return mlp_forward(sd, head, mlp_forward(sd, enc, obs.float()))


def critic_forward(sd, specs, obs: torch.Tensor, act: torch.Tensor) -> torch.Tensor:
    enc, head = specs
    latent = mlp_forward(sd, enc, obs.float())
    return mlp_forward(sd, head, torch.cat([latent, act], dim=-1))        # q_networks.py:424-425


def _leaf(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


class OracleDDPG:
    """State = reference state_dicts of actor / critic(s) and their targets; ``twin=True`` is TD3."""

    def __init__(self, a_specs, c_specs, actor_sd, actor_target_sd, critic_sds, critic_target_sds, *, gamma=0.99,
                 tau=1e-3, lr_actor=1e-4, lr_critic=1e-3, policy_freq=2, action_low=-1.0, action_high=1.0,
                 twin: bool = False):
        assert len(critic_sds) == (2 if twin else 1)
        self.a_specs, self.c_specs, self.twin = a_specs, c_specs, twin
        self.actor, self.actor_target = _leaf(actor_sd), {k: v.clone() for k, v in actor_target_sd.items()}
        self.critics = [_leaf(sd) for sd in critic_sds]
        self.critic_targets = [{k: v.clone() for k, v in sd.items()} for sd in critic_target_sds]
        self.gamma, self.tau, self.policy_freq = gamma, tau, policy_freq
        self.low, self.high = action_low, action_high
        self.opt_actor = torch.optim.Adam(list(self.actor.values()), lr=lr_actor)
        self.opt_critics = [torch.optim.Adam(list(c.values()), lr=lr_critic) for c in self.critics]
        self.learn_counter = 0
        self.last_grads: dict = {}

    def _soft(self, net, target):
        """ddpg.py:496-508: target = tau*param + (1-tau)*target over parameters()."""
        with torch.no_grad():
            for k in net:
                target[k].copy_(self.tau * net[k].data + (1.0 - self.tau) * target[k])

    def learn(self, exp: dict, noise_clip: float = 0.5, policy_noise: float = 0.2):
        obs, actions, rewards, next_obs, dones = (exp[k] for k in ("obs", "action", "reward", "next_obs", "done"))
        q = [critic_forward(c, self.c_specs, obs, actions) for c in self.critics]
        with torch.no_grad():
            next_actions = actor_forward(self.actor_target, self.a_specs, next_obs)
            noise = actions.data.normal_(0, policy_noise)                  # in place on the batch (quirk)
            noise = torch.clamp(noise, -noise_clip, noise_clip)
            next_actions = torch.clamp(next_actions + noise, self.low, self.high)
            qn = [critic_forward(t, self.c_specs, next_obs, next_actions) for t in self.critic_targets]
            q_next = torch.min(qn[0], qn[1]) if self.twin else qn[0]
        y = rewards + ((1 - dones) * self.gamma * q_next)
        critic_loss = F.mse_loss(q[0], y)
        if self.twin:
            critic_loss = critic_loss + F.mse_loss(q[1], y)
        for o in self.opt_critics:
            o.zero_grad()
        critic_loss.backward()
        self.last_grads = {f"critic{i}/{k}": v.grad.detach().clone() for i, c in enumerate(self.critics)
                           for k, v in c.items()}
        for o in self.opt_critics:
            o.step()
        self.learn_counter += 1
        actor_loss = None
        if self.learn_counter % self.policy_freq == 0:
            policy_actions = actor_forward(self.actor, self.a_specs, obs)
            a_loss = -critic_forward(self.critics[0], self.c_specs, obs, policy_actions).mean()
            self.opt_actor.zero_grad()
            a_loss.backward()
            self.last_grads.update({f"actor/{k}": v.grad.detach().clone() for k, v in self.actor.items()})
            self.opt_actor.step()
            self._soft(self.actor, self.actor_target)
            for c, t in zip(self.critics, self.critic_targets):
                self._soft(c, t)
            actor_loss = a_loss.item()
        return actor_loss, critic_loss.item()
