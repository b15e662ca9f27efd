"""GPU (needs >= 2 devices): population sharded one agent per GPU — NCCL fitness all-gather and
device-to-device move of the tournament winner's flat buffers."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from agilerl_b200.algorithms import RainbowDQN
    from agilerl_b200.compat import spaces
    from agilerl_b200.hpo import TournamentSelection
    net = {"encoder_config": {"channel_size": [8, 16], "kernel_size": [4, 3], "stride_size": [2, 1]},
           "head_config": {"hidden_size": [32]}, "latent_dim": 16}
    torch.manual_seed(100 + rank)
    agent = RainbowDQN(spaces.Box(0, 255, (3, 20, 20), np.uint8), spaces.Discrete(4), index=rank, net_config=net,
                       batch_size=8, v_min=-10.0, v_max=10.0, device=f"cuda:{rank}")
    agent.fitness = [10.0 if rank == 1 else 1.0]          # rank 1 owns the elite
    if rank == 1:                                          # give it a different architecture + optimiser state
        agent.actor.encoder.add_channel(hidden_layer=0, numb_new_channels=8)
        agent.actor_target = type(agent.actor)(**agent.actor.init_dict)
        agent.actor_target.load_state_dict(agent.actor.state_dict())
        agent.reinit_optimizers()
        agent.engine.exp_avg.fill_(0.25); agent.engine.step = 5
    before = agent.actor.buffers.params.sum().item()
    ts = TournamentSelection(2, True, world, 1, seed=3)
    elite, new_pop = ts.select([agent])
    child = new_pop[0]
    obs = np.zeros((2, 3, 20, 20), np.uint8)
    out[rank] = dict(plan=ts.last_plan, index=child.index, channels=list(child.actor.encoder.channel_size),
                     psum=child.actor.buffers.params.sum().item(), before=before,
                     exp_avg=child.engine.exp_avg.mean().item(), step=child.engine.step,
                     act=child.get_action(obs).tolist(), fitness=child.fitness)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_sharded_tournament_moves_winner():
    import torch.multiprocessing as mp
    world = 2
    port = 29600 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["plan"] == r1["plan"]
    elite_pos, slots = r0["plan"]
    assert elite_pos == 1 and slots[0] == (1, 1)           # slot 0 (rank 0) receives the elite from rank 1
    assert r0["index"] == 1 and r0["channels"] == [16, 16] and r0["fitness"] == [10.0]
    assert abs(r0["psum"] - r1["before"]) < 1e-3           # rank 0 now holds rank 1's weights
    assert abs(r0["exp_avg"] - 0.25) < 1e-6 and r0["step"] == 5   # ... and its Adam state
    assert r1["index"] == 2                                 # slot 1: fresh index max_id + 1
    assert len(r0["act"]) == 2 and len(r1["act"]) == 2
