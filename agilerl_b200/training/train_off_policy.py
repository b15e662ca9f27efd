"""``train_off_policy`` — same signature and control flow as
agilerl/training/train_off_policy.py:41-616 (the north star keeps the reference's file as the
driver; it cannot be imported in the build image — gymnasium / accelerate / tensordict / wandb are
absent — so this module restates its loop for the algorithms of this path, minus W&B / accelerate /
checkpoint plumbing).  One addition: ``fused=True`` routes PER + n-step learning through the
HBM-resident fused step (``RainbowDQN.learn_from_buffers``) instead of sample -> learn ->
update_priorities through host-visible tensors; ``share_experience=True`` (population sharded one process per GPU)
all-gathers every environment step across the ranks before ingest, so each rank's buffer holds the whole
population's experience like the reference's single shared buffer (``training.population.share_transitions``).
"""
from __future__ import annotations

import time
import warnings

import numpy as np
import torch

from ..algorithms import DDPG, DQN, TD3, RainbowDQN
from ..algorithms.dqn_rainbow import obs_channels_to_first
from ..networks.actors import DeterministicActor
from ..components import MultiStepReplayBuffer, PrioritizedReplayBuffer, ReplayBuffer, Sampler, Transition
from ..utils.utils import tournament_selection_and_mutation
from .population import share_transitions


def train_off_policy(env, env_name: str, algo: str, pop: list, memory: ReplayBuffer, INIT_HP: dict | None = None,
                     MUT_P: dict | None = None, swap_channels: bool = False, max_steps: int = 1000000,
                     evo_steps: int = 10000, eval_steps: int | None = None, eval_loop: int = 1,
                     learning_delay: int = 0, eps_start: float = 1.0, eps_end: float = 0.1, eps_decay: float = 0.995,
                     target: float | None = None, n_step: bool = False, per: bool = False,
                     n_step_memory: MultiStepReplayBuffer | None = None, tournament=None, mutation=None,
                     checkpoint: int | None = None, checkpoint_path: str | None = None, overwrite_checkpoints: bool = False,
                     save_elite: bool = False, elite_path: str | None = None, wb: bool = False, verbose: bool = True,
                     accelerator=None, wandb_api_key: str | None = None, wandb_kwargs: dict | None = None,
                     fused: bool = False, share_experience: bool = False):
    assert isinstance(algo, str), "'algo' must be the name of the algorithm as a string."
    assert isinstance(max_steps, int), "Number of steps must be an integer."
    assert isinstance(evo_steps, int), "Evolution frequency must be an integer."
    assert accelerator is None, "accelerate is replaced by one-agent-per-GPU sharding in agilerl_b200"
    if n_step:
        assert isinstance(n_step_memory, MultiStepReplayBuffer), "n_step_memory must be a MultiStepReplayBuffer"
    if per:
        assert isinstance(memory, PrioritizedReplayBuffer), "memory must be a PrioritizedReplayBuffer when per=True"
    if wb:
        warnings.warn("W&B logging is outside the hot path and is not wired here", stacklevel=2)
    if fused:
        assert per and n_step_memory is not None, "the fused step needs PER + n-step buffers"
    num_envs = env.num_envs if hasattr(env, "num_envs") else 1
    is_vectorised = hasattr(env, "num_envs")
    sampler = Sampler(memory=memory)                                   # train_off_policy.py:220-222
    n_step_sampler = Sampler(memory=n_step_memory) if n_step_memory is not None else None
    pop_loss = [[] for _ in pop]
    pop_fitnesses = []
    total_steps = 0
    loss = None
    if mutation is not None:
        pop = mutation.mutation(pop, pre_training_mut=True)            # :238-240

    def learn_once(agent):
        if fused:
            return agent.learn_from_buffers(memory, n_step_memory)     # device tensor, no sync
        if per:
            experiences = sampler.sample(agent.batch_size, agent.beta)
            n_exp = n_step_sampler.sample(experiences["idxs"]) if n_step_memory is not None else None
            loss_, idxs, priorities = agent.learn(experiences, n_experiences=n_exp, per=per)
            memory.update_priorities(idxs, priorities)
            return loss_
        experiences = sampler.sample(agent.batch_size, return_idx=n_step_memory is not None)
        if n_step_memory is not None:
            loss_, *_ = agent.learn(experiences, n_experiences=n_step_sampler.sample(experiences["idxs"]))
            return loss_
        loss_ = agent.learn(experiences)
        return loss_[0] if isinstance(agent, RainbowDQN) else loss_

    while np.less([agent.steps[-1] for agent in pop], max_steps).all():
        pop_episode_scores, pop_fps = [], []
        for agent_idx, agent in enumerate(pop):
            obs, info = env.reset()
            scores = np.zeros(num_envs)
            completed_episode_scores, losses = [], []
            steps = 0
            epsilon = eps_start
            start_time = time.time()
            for idx_step in range(evo_steps // num_envs):
                if swap_channels:
                    obs = obs_channels_to_first(obs)
                if isinstance(agent, DQN):
                    action = agent.get_action(obs, epsilon, action_mask=info.get("action_mask", None))
                    epsilon = max(eps_end, epsilon * eps_decay)
                elif isinstance(agent, RainbowDQN):
                    action = agent.get_action(obs, action_mask=info.get("action_mask", None))
                else:                                                   # DDPG / TD3 (:281-293): the environment gets the
                    raw_action = agent.get_action(obs)                  # action rescaled to its bounds, the buffer the raw one
                    action = DeterministicActor.rescale_action(action=torch.from_numpy(raw_action), low=agent.action_low,
                                                               high=agent.action_high,
                                                               output_activation=agent.actor.output_activation).cpu().numpy()
                if not is_vectorised:
                    action = action[0]
                next_obs, reward, done, trunc, info = env.step(action)
                scores += np.array(reward)
                if not is_vectorised:
                    done, trunc = np.array([done]), np.array([trunc])
                reset_noise_indices = []
                for idx, (d, t) in enumerate(zip(done, trunc)):
                    if d or t:
                        completed_episode_scores.append(scores[idx])
                        agent.scores.append(scores[idx])
                        scores[idx] = 0
                        reset_noise_indices.append(idx)
                if isinstance(agent, (DDPG, TD3)):                      # :312-313, :325-326
                    agent.reset_action_noise(reset_noise_indices)
                    action = raw_action
                total_steps += num_envs
                steps += num_envs
                next_obs = obs_channels_to_first(next_obs) if swap_channels else next_obs
                if is_vectorised:
                    transition = Transition(obs=obs, action=action, reward=reward, next_obs=next_obs, done=done,
                                            batch_size=[num_envs]).to_tensordict()
                else:
                    transition = Transition(obs=np.asarray(obs)[None], action=np.asarray([action]),
                                            reward=np.asarray([reward]), next_obs=np.asarray(next_obs)[None],
                                            done=done, batch_size=[1]).to_tensordict()
                if share_experience:                                    # sharded population: every rank ingests every rank's steps
                    transition = share_transitions(transition, getattr(memory, "_dev", None))
                if n_step_memory is not None:
                    one_step = n_step_memory.add(transition)
                    if one_step is not None:
                        memory.add(one_step)
                else:
                    memory.add(transition)
                if per:                                                 # :346-351
                    fraction = min(((agent.steps[-1] + idx_step + 1) * num_envs / max_steps), 1.0)
                    agent.beta += fraction * (1.0 - agent.beta)
                ready = len(memory) >= agent.batch_size and memory.size > learning_delay
                if agent.learn_step > num_envs:
                    if idx_step % (agent.learn_step // num_envs) == 0 and ready:
                        loss = learn_once(agent)
                elif ready:
                    for _ in range(num_envs // agent.learn_step):
                        loss = learn_once(agent)
                if loss is not None:
                    losses.append(loss)
                obs = next_obs
            agent.steps[-1] += steps
            pop_fps.append(steps / max(time.time() - start_time, 1e-12))
            pop_episode_scores.append(completed_episode_scores)
            if losses:
                if isinstance(losses[-1], tuple):          # DDPG / TD3: (actor_loss | None, critic_loss) — train_off_policy.py:445-453
                    actor_losses, critic_losses = list(zip(*losses))
                    pop_loss[agent_idx].append((np.mean([l for l in actor_losses if l is not None]), np.mean(critic_losses)))
                else:
                    vals = [float(l.item()) if hasattr(l, "item") else float(l) for l in losses]
                    pop_loss[agent_idx].append(np.mean(vals))
        if isinstance(agent, DQN):        # train_off_policy.py:458-460: ONCE per generation, from the last agent's epsilon
            eps_start = epsilon
        fitnesses = [agent.test(env, swap_channels=swap_channels, max_steps=eval_steps, loop=eval_loop) for agent in pop]
        pop_fitnesses.append(fitnesses)
        if verbose:
            print(f"--- Global steps {total_steps} --- fitness {['%.2f' % f for f in fitnesses]} "
                  f"fps {['%.0f' % f for f in pop_fps]} mutations {[a.mut for a in pop]}")
        for agent in pop:
            agent.steps.append(agent.steps[-1])
        if target is not None and np.all(np.greater([np.mean(a.fitness[-10:]) for a in pop], target)) and \
                len(pop_fitnesses) >= 10:
            break
        if tournament and mutation is not None:
            pop = tournament_selection_and_mutation(population=pop, tournament=tournament, mutation=mutation,
                                                    env_name=env_name, algo=algo, elite_path=elite_path,
                                                    save_elite=save_elite)
    return pop, pop_fitnesses
