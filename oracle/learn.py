"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of the reference's ``learn()`` for Rainbow-DQN and DQN: C51 categorical
projection loss, PER weighting, gradient clipping, Adam, Polyak update and noise reset.  Uses the
same torch CPU ops the reference dispatches (the arithmetic lives in PyTorch).

Follows (reference file:line):
  * RainbowDQN._dqn_loss        agilerl/algorithms/dqn_rainbow.py:284-367
  * RainbowDQN.learn            agilerl/algorithms/dqn_rainbow.py:369-490
  * RainbowDQN.soft_update      agilerl/algorithms/dqn_rainbow.py:492-501
  * DQN.update / learn          agilerl/algorithms/dqn.py:274-358
  * torch.optim.Adam defaults   (betas 0.9/0.999, eps 1e-8) via OptimizerWrapper,
                                agilerl/algorithms/core/optimizer_wrapper.py:338-356

Pinned against the unmodified reference run through ``oracle.refshim``
(tests/golden/make_golden.py, tests/test_oracle_vs_reference.py).
"""
from __future__ import annotations

import numpy as np
import torch
from torch.nn.utils import clip_grad_norm_

from . import nets


class OracleAgent:
    """State of one agent: online/target state-dicts, Adam moments, hyper-parameters."""

    def __init__(self, spec: nets.NetSpec, actor_sd: dict, target_sd: dict | None = None, *,
                 batch_size=64, lr=1e-4, gamma=0.99, tau=1e-3, prior_eps=1e-6, v_min=-10.0,
                 v_max=10.0, n_step=3, combined_reward=False, double=False):
        self.spec = spec
        self.actor = {k: v.detach().clone() for k, v in actor_sd.items()}
        src = actor_sd if target_sd is None else target_sd
        self.target = {k: v.detach().clone() for k, v in src.items()}
        self.pkeys = nets.param_keys(self.actor)
        for k in self.pkeys:
            self.actor[k].requires_grad_(True)
        self.batch_size, self.lr, self.gamma, self.tau = batch_size, lr, gamma, tau
        self.prior_eps, self.v_min, self.v_max, self.n_step = prior_eps, v_min, v_max, n_step
        self.combined_reward, self.double = combined_reward, double
        if spec.kind == "rainbow":
            # dqn_rainbow.py:165-171
            self.support = torch.linspace(v_min, v_max, spec.num_atoms)
            self.delta_z = (v_max - v_min) / (spec.num_atoms - 1)
            spec.support = self.support
        self.opt = torch.optim.Adam([self.actor[k] for k in self.pkeys], lr=lr)

    # ---- Rainbow --------------------------------------------------------------------------
    def dqn_loss(self, obs, actions, rewards, next_obs, dones, gamma) -> torch.Tensor:
        """dqn_rainbow.py:284-367, literally (including the broadcasting behaviour that the
        shapes of ``rewards``/``dones`` induce — quirk Q2)."""
        spec, B, N = self.spec, self.batch_size, self.spec.num_atoms
        obs = nets.preprocess(spec, obs)
        next_obs = nets.preprocess(spec, next_obs)
        with torch.no_grad():
            next_actions = nets.rainbow_forward(self.actor, spec, next_obs).argmax(1)
            target_q_dist = nets.rainbow_forward(self.target, spec, next_obs, q=False)
            target_q_dist = target_q_dist[range(B), next_actions]
            t_z = rewards + (1 - dones) * gamma * self.support
            t_z = t_z.clamp(min=self.v_min, max=self.v_max)
            b = (t_z - self.v_min) / self.delta_z
            L = b.floor().long()
            u = b.ceil().long()
            L[(u > 0) * (u == L)] -= 1
            u[((N - 1) > L) * (u == L)] += 1
            offset = (torch.linspace(0, (B - 1) * N, B).long().unsqueeze(1).expand(B, N))
            proj_dist = torch.zeros(target_q_dist.size())
            proj_dist.view(-1).index_add_(0, (L + offset).view(-1),
                                          (target_q_dist * (u.float() - b)).view(-1))
            proj_dist.view(-1).index_add_(0, (u + offset).view(-1),
                                          (target_q_dist * (b - L.float())).view(-1))
        log_q_dist = nets.rainbow_forward(self.actor, spec, obs, q=False, log=True)
        log_p = log_q_dist[range(B), actions.squeeze().long()]
        self.last_proj_dist = proj_dist
        return -(proj_dist * log_p).sum(1)

    def learn_rainbow(self, experiences: dict, n_experiences: dict | None = None, per: bool = False,
                      noise_normals: tuple[torch.Tensor, torch.Tensor] | None = None):
        """dqn_rainbow.py:369-490.  ``noise_normals`` = (actor, target) flat standard-normal
        vectors consumed by the two ``reset_noise`` calls (:484-485); ``None`` draws them from
        the global torch RNG in the reference's order."""
        n_step = n_experiences is not None
        e = experiences
        elementwise = None
        if self.combined_reward or not n_step:
            elementwise = self.dqn_loss(e["obs"], e["action"], e["reward"], e["next_obs"], e["done"],
                                        self.gamma)
        if n_step:
            n = n_experiences
            n_loss = self.dqn_loss(n["obs"], n["action"], n["reward"], n["next_obs"], n["done"],
                                   self.gamma ** self.n_step)
            elementwise = elementwise + n_loss if self.combined_reward else n_loss
        if per:
            idxs = e["idxs"]
            loss = torch.mean(elementwise * e["weights"])      # quirk Q1: broadcasting kept
        else:
            idxs = e["idxs"] if n_step else None
            loss = torch.mean(elementwise)
        self.opt.zero_grad()
        loss.backward()
        self.last_grads = {k: self.actor[k].grad.detach().clone() for k in self.pkeys}
        clip_grad_norm_([self.actor[k] for k in self.pkeys], 10.0)
        self.opt.step()
        self.soft_update()
        self.reset_noise(noise_normals)
        new_priorities = None
        if per:
            new_priorities = elementwise.detach().cpu().numpy() + self.prior_eps
        return loss.item(), idxs, new_priorities

    def soft_update(self) -> None:
        """dqn_rainbow.py:492-501 / dqn.py:349-358."""
        with torch.no_grad():
            for k in self.pkeys:
                self.target[k].copy_(self.tau * self.actor[k] + (1.0 - self.tau) * self.target[k])

    def reset_noise(self, noise_normals=None) -> None:
        with torch.no_grad():
            for sd, i in ((self.actor, 0), (self.target, 1)):
                n = nets.num_noise_normals(self.spec)
                if n == 0:
                    continue
                if noise_normals is None:
                    # the reference draws randn(in) then randn(out) per layer, one call each
                    # (custom_components.py:118-119,130): keep the same RNG stream order
                    z = torch.cat([torch.cat([torch.randn(a), torch.randn(b)])
                                   for _, a, b in nets.noisy_layer_keys(self.spec)])
                else:
                    z = noise_normals[i]
                nets.reset_noise_from_normals(sd, self.spec, z)

    # ---- DQN ------------------------------------------------------------------------------
    def learn_dqn(self, experiences: dict) -> float:
        """dqn.py:326-347 + update :274-324 (MSE, no clipping, Adam, soft update)."""
        spec = self.spec
        obs = nets.preprocess(spec, experiences["obs"])
        next_obs = nets.preprocess(spec, experiences["next_obs"])
        actions, rewards, dones = experiences["action"], experiences["reward"], experiences["done"]
        with torch.no_grad():
            if self.double:
                q_idx = nets.q_forward(self.actor, spec, next_obs).argmax(dim=1).unsqueeze(1)
                q_target = nets.q_forward(self.target, spec, next_obs).gather(dim=1, index=q_idx)
            else:
                q_target = nets.q_forward(self.target, spec, next_obs).max(axis=1)[0].unsqueeze(1)
            y_j = rewards + self.gamma * q_target * (1 - dones)
        if actions.ndim == 1:
            actions = actions.unsqueeze(-1)
        q_eval = nets.q_forward(self.actor, spec, obs).gather(1, actions.long())
        loss = torch.nn.functional.mse_loss(q_eval, y_j)
        self.opt.zero_grad()
        loss.backward()
        self.last_grads = {k: self.actor[k].grad.detach().clone() for k in self.pkeys}
        self.opt.step()
        self.soft_update()
        return loss.item()

    # ---- acting ---------------------------------------------------------------------------
    def q_values(self, obs, training: bool = True) -> torch.Tensor:
        """get_action's forward (dqn_rainbow.py:258-262 / dqn.py:262-264)."""
        x = nets.preprocess(self.spec, obs)
        with torch.no_grad():
            if self.spec.kind == "rainbow":
                return nets.rainbow_forward(self.actor, self.spec, x, train_noise=training)
            return nets.q_forward(self.actor, self.spec, x)


def c51_projection(target_q_dist, rewards, dones, gamma, support, v_min, v_max, delta_z):
    """Stand-alone restatement of dqn_rainbow.py:323-360 for canonical ``[B,1]`` reward/done.
    Returns (proj_dist [B,N], L, u, w_l=(u-b), w_u=(b-L))."""
    B, N = target_q_dist.shape
    t_z = (rewards + (1 - dones) * gamma * support).clamp(min=v_min, max=v_max)
    b = (t_z - v_min) / delta_z
    L = b.floor().long()
    u = b.ceil().long()
    L[(u > 0) * (u == L)] -= 1
    u[((N - 1) > L) * (u == L)] += 1
    offset = torch.linspace(0, (B - 1) * N, B).long().unsqueeze(1).expand(B, N)
    proj = torch.zeros(B, N)
    w_l, w_u = u.float() - b, b - L.float()
    proj.view(-1).index_add_(0, (L + offset).view(-1), (target_q_dist * w_l).view(-1))
    proj.view(-1).index_add_(0, (u + offset).view(-1), (target_q_dist * w_u).view(-1))
    return proj, L, u, w_l, w_u


def to_numpy(sd: dict) -> dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}
