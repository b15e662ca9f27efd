"""Parity of the BENCHMARKED path at BASELINE config 2's real size.

``LearnEngine.rainbow_fused_step`` — the call ``bench.py``'s ``value`` times — with the north-star network
(4x84x84 uint8, conv 32/32 k8/4 s4/2, latent 32, noisy dueling head [64], 6 actions, 51 atoms), **B = 256**,
frames read through ``row_idx`` from a ring of 160 000 rows (4.5 GB per frame tensor: byte offsets beyond
2^32; the sampled rows are steered to the top of the ring), trees of capacity 2^18, against ``oracle.replay``
+ ``oracle.learn`` on the same rows (reference: dqn_rainbow.py:284-490, replay_buffer.py:331-428).

Bars: sampled indices, IS weights' inputs and every tree node bit-exact; loss / priorities within 1e-5;
every gradient tensor within 2e-5 max|g|.
"""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROWS = 160_000                       # 160000 * 28224 B = 4.52 GB > 2^32
TOP = 2 ** 32 // 28224 + 1           # first row whose byte offset exceeds 2^32
OBS, A, B = (4, 84, 84), 6, 256
ALPHA, BETA, GAMMA, NSTEP = 0.6, 0.4, 0.99, 3
HP = dict(v_min=-10.0, v_max=10.0, delta_z=20.0 / 50, lr=1e-4, tau=1e-3, prior_eps=1e-6)


def _state(layout, gen):
    sd = {}
    for k, e in layout.entries.items():
        fan = e.shape[-1] if len(e.shape) > 1 else e.shape[0]
        if "norm" in k and k.endswith("weight"):
            sd[k] = torch.ones(e.shape) + 0.1 * torch.randn(e.shape, generator=gen)
        elif "epsilon" in k:
            sd[k] = torch.randn(e.shape, generator=gen) * 0.5
        elif "conv" in k and k.endswith("weight"):
            sd[k] = torch.randn(e.shape, generator=gen) * (1.0 / (e.shape[1] * e.shape[2] * e.shape[3]) ** 0.5)
        else:
            sd[k] = torch.randn(e.shape, generator=gen) * (0.5 / fan ** 0.5)
    return sd


@pytest.fixture(scope="module")
def ring():
    """Replay pair filled like bench.py's build_rank (already-rolled n-step transitions), 160k rows, with the
    priorities of the rows above 2^32 bytes boosted so that about half of every batch lands there; the
    oracle's trees hold the same leaves (host ``p ** alpha``, bit-identical)."""
    from agilerl_b200.compat import TensorDict
    from agilerl_b200.components import MultiStepReplayBuffer, PrioritizedReplayBuffer
    from agilerl_b200.components.replay_buffer import ReplayBuffer
    from oracle import replay as oreplay
    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("needs 24 GB of free HBM")
    dev = "cuda"
    mem, nmem = PrioritizedReplayBuffer(ROWS, ALPHA, device=dev), MultiStepReplayBuffer(ROWS, NSTEP, GAMMA, device=dev)
    g = torch.Generator(device=dev).manual_seed(2024)
    for s in range(0, ROWS, 10_000):
        n = min(10_000, ROWS - s)
        small = dict(action=torch.randint(0, A, (n,), device=dev, generator=g).float(),
                     reward=torch.randn(n, device=dev, generator=g),
                     done=(torch.rand(n, device=dev, generator=g) < 0.05).float())
        td = TensorDict(dict(small, obs=torch.randint(0, 256, (n, *OBS), dtype=torch.uint8, device=dev, generator=g),
                             next_obs=torch.randint(0, 256, (n, *OBS), dtype=torch.uint8, device=dev, generator=g)),
                        batch_size=[n])
        ReplayBuffer.add(nmem, td)
        nmem.done_key = "done"
        mem.add(TensorDict(small, batch_size=[n]))      # the PER buffer's own frames are never read on this path
    cg = torch.Generator().manual_seed(7)
    pri = torch.randn(ROWS, generator=cg).abs() + 1e-6
    pri[TOP:] *= (TOP / (ROWS - TOP)) ** (1.0 / ALPHA)  # equal p**alpha mass below and above the 2^32-byte line
    pri = pri.numpy().astype(np.float32)
    omem = oreplay.OraclePER(ROWS, ALPHA)
    omem.size = ROWS
    for s in range(0, ROWS, 8192):
        sl = slice(s, min(s + 8192, ROWS))
        mem.update_priorities(torch.arange(sl.start, sl.stop), pri[sl])
        omem.update_priorities(range(sl.start, sl.stop), pri[sl])
    torch.cuda.synchronize()
    assert nmem._fields[("obs",)].numel() > 2 ** 32
    return mem, nmem, omem


def _trees(mem):
    return mem.sum_tree._t.cpu().numpy().copy(), mem.min_tree._t.cpu().numpy().copy()


def _engine(seed):
    from agilerl_b200.engine import LearnEngine, NetBuffers
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    from oracle import learn as olearn, nets as onets
    layout = FlatLayout(rainbow_spec(OBS, A, channel_size=(32, 32), kernel_size=(8, 4), stride_size=(4, 2),
                                     obs_low=0.0, obs_high=255.0, obs_u8=True))
    assert layout.n_param_elems == 162730
    gen = torch.Generator().manual_seed(seed)
    sd_a = _state(layout, gen)
    sd_t = {k: v + 0.01 * torch.randn(v.shape, generator=gen) for k, v in sd_a.items()}
    actor, target = NetBuffers(layout, "cuda"), NetBuffers(layout, "cuda")
    actor.load_state_dict(sd_a); target.load_state_dict(sd_t)
    eng = LearnEngine(layout, actor, target)
    oa = olearn.OracleAgent(onets.rainbow_spec(OBS, A), sd_a, sd_t, batch_size=B, lr=HP["lr"], v_min=-10.0, v_max=10.0)
    return eng, layout, oa, gen


def _oracle_batch(nmem, idx_cpu, weights):
    f = nmem._fields
    take = lambda key: f[(key,)][idx_cpu.cuda()].cpu()
    nexp = dict(obs=take("obs"), action=take("action"), reward=take("reward"), next_obs=take("next_obs"), done=take("done"))
    exp = dict(nexp, weights=weights, idxs=idx_cpu)
    return exp, nexp


def _check_trees_against_leaves(mem, idx, pri_dev):
    """Leaves written on device = pow(max(p, 1e-5), alpha) of the device's own priorities (<= 1 ulp from
    CPython's); every internal node bit-equal to op(children) recomputed from those leaves."""
    st, mt = _trees(mem)
    cap = mem._cap
    p = pri_dev.cpu().numpy().astype(np.float64)
    want = np.array([max(float(x), 1e-5) ** ALPHA for x in p])
    got = st[cap + idx.cpu().numpy()]
    last = {}
    for j, i in enumerate(idx.cpu().numpy()):
        last[int(i)] = want[j]                           # duplicates: last writer wins
    for i, w in last.items():
        assert abs(st[cap + i] - w) <= np.spacing(w), (i, st[cap + i], w)
        assert mt[cap + i] == st[cap + i]
    lvl_s, lvl_m = st[cap:2 * cap].copy(), mt[cap:2 * cap].copy()
    n = cap
    while n > 1:
        lvl_s = lvl_s[0::2] + lvl_s[1::2]
        lvl_m = np.minimum(lvl_m[0::2], lvl_m[1::2])
        n //= 2
        np.testing.assert_array_equal(st[n:2 * n], lvl_s)
        np.testing.assert_array_equal(mt[n:2 * n], lvl_m)
    return got


def _compare_step(eng, layout, oa, loss, pri, oloss, opri, check_grads=True):
    from agilerl_b200.engine import NetBuffers
    assert abs(loss.item() - oloss) <= 1e-5 * max(1.0, abs(oloss)), (loss.item(), oloss)
    np.testing.assert_allclose(pri.cpu().numpy(), np.asarray(opri).reshape(-1), rtol=1e-5, atol=1e-5)
    if not check_grads:
        return
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in oa.last_grads.values())).item()
    coef = min(1.0, 10.0 / (total + 1e-6))               # clip_grad_norm_ scales the stored gradients
    gv = NetBuffers(layout, "cuda"); gv.params.copy_(eng.grads)
    for k, ref in oa.last_grads.items():
        ref = ref * coef
        scale = max(ref.abs().max().item(), 1e-6)
        err = (gv.view(k).cpu() - ref).abs().max().item()
        assert err <= 2e-5 * scale + 1e-8, f"grad {k}: {err} vs max|g| {scale}"


def _fused(eng, mem, nmem, u, z):
    support = torch.linspace(-10.0, 10.0, 51).cuda()
    return eng.rainbow_fused_step(mem, nmem, B=B, beta=BETA, support=support, hp=HP, gamma_n=GAMMA ** NSTEP,
                                  uniforms=None if u is None else u.cuda(), noise_normals=z)


def test_fused_step_full_size_injected_randomness(ring):
    mem, nmem, omem = ring
    eng, layout, oa, gen = _engine(11)
    u = torch.rand(B, generator=gen)
    z = (torch.randn(eng.noise_count, generator=gen), torch.randn(eng.noise_count, generator=gen))
    st0, mt0 = _trees(mem)
    np.testing.assert_array_equal(st0, omem.sum_tree.tree)        # same leaves, same sums before the step
    np.testing.assert_array_equal(mt0, omem.min_tree.tree)
    a0 = {k: eng.actor.view(k).cpu().clone() for k in oa.pkeys}
    loss, idx, pri = _fused(eng, mem, nmem, u, z)
    torch.cuda.synchronize()
    oidx = omem.sample_proportional(B, uniforms=u)
    assert torch.equal(idx.cpu(), oidx), "sampled indices differ from the oracle"
    assert int((oidx >= TOP).sum()) >= B // 4, "too few rows beyond the 2^32-byte line for this test to mean anything"
    ow = omem.calculate_weights(oidx, BETA)
    exp, nexp = _oracle_batch(nmem, oidx, ow)
    oloss, _, opri = oa.learn_rainbow(exp, nexp, per=True, noise_normals=z)
    _compare_step(eng, layout, oa, loss, pri, oloss, opri)
    # optimiser result against the oracle's parameters where Adam's first step has a clear sign
    for k in oa.pkeys:
        p1, ref = eng.actor.view(k).cpu(), oa.actor[k].detach()
        mask = oa.last_grads[k].abs() > 1e-5
        if mask.any():
            assert (p1 - ref)[mask].abs().max().item() <= 5e-6, f"param {k}"
        assert not torch.equal(p1, a0[k]) or not mask.any()
    _check_trees_against_leaves(mem, idx, pri)
    # the oracle's write-back of ITS priorities: leaves agree to the 1e-5 the priorities agree to
    omem.update_priorities(oidx, opri)
    st1, _ = _trees(mem)
    cap = mem._cap
    np.testing.assert_allclose(st1[cap + oidx.numpy()], omem.sum_tree.tree[cap + oidx.numpy()], rtol=2e-5)
    np.testing.assert_allclose(st1[1], omem.sum_tree.tree[1], rtol=1e-9)


def test_fused_step_full_size_philox_streams_read_back(ring):
    """Production randomness: uniforms and NoisyLinear normals drawn on device (Philox); the same streams are
    read back through b2rl_philox_uniforms / b2rl_philox_normals and handed to the oracle."""
    from agilerl_b200 import _lib
    mem, nmem, omem = ring
    # bring the oracle's trees to the device's current state (earlier tests wrote device-pow leaves)
    st, mt = _trees(mem)
    omem.sum_tree.tree[:] = st
    omem.min_tree.tree[:] = mt
    eng, layout, oa, gen = _engine(12)
    z0 = (torch.randn(eng.noise_count, generator=gen), torch.randn(eng.noise_count, generator=gen))
    eng.reset_noise(eng.actor, z0[0]); eng.reset_noise(eng.target, z0[1])
    from oracle import nets as onets
    onets.reset_noise_from_normals(oa.actor, oa.spec, z0[0]); onets.reset_noise_from_normals(oa.target, oa.spec, z0[1])
    lib, s = _lib.load(), _lib.stream_ptr(torch.device("cuda", torch.cuda.current_device()))
    u = torch.empty(B, device="cuda")
    _lib.check(lib.b2rl_philox_uniforms(mem._philox_seed, mem._philox_offset, B, u.data_ptr(), s))
    nz = eng.noise_count
    z = torch.empty(2 * nz, device="cuda")
    _lib.check(lib.b2rl_philox_normals(eng.philox_seed, eng.philox_offset, 2 * nz, z.data_ptr(), s))
    was = mem.device_rng
    mem.device_rng = True
    try:
        loss, idx, pri = _fused(eng, mem, nmem, None, None)
    finally:
        mem.device_rng = was
    torch.cuda.synchronize()
    u, z = u.cpu(), z.cpu()
    assert 0.0 <= float(u.min()) and float(u.max()) < 1.0 and abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1) < 0.05
    oidx = omem.sample_proportional(B, uniforms=u)
    assert torch.equal(idx.cpu(), oidx)
    exp, nexp = _oracle_batch(nmem, oidx, omem.calculate_weights(oidx, BETA))
    oloss, _, opri = oa.learn_rainbow(exp, nexp, per=True, noise_normals=(z[:nz], z[nz:]))
    _compare_step(eng, layout, oa, loss, pri, oloss, opri)
    for k, v in oa.actor.items():                       # the epsilon buffers the device drew == f(z) of the read-back normals
        if k.endswith("_epsilon"):
            np.testing.assert_allclose(eng.actor.view(k).cpu().numpy(), v.numpy(), rtol=2.5e-7, atol=0)
            np.testing.assert_allclose(eng.target.view(k).cpu().numpy(), oa.target[k].numpy(), rtol=2.5e-7, atol=0)
    _check_trees_against_leaves(mem, idx, pri)


def test_twenty_fused_steps_device_pow_writeback_do_not_drift(ring):
    """20 consecutive fused steps (device-``pow`` leaves from the device's own float32 priorities) beside 20
    oracle steps (host ``**`` on the oracle's priorities).  The two sides' leaves differ by what the
    priorities differ by (<= 1e-5 relative), so a stratified draw can only change if its upper bound falls
    within that distance of a leaf boundary.  Stated bar: the sampled indices are IDENTICAL on every one of
    the 20 steps (5120 draws) for this seed, losses stay within 1e-5 (first step) / 1e-4 (every step) of the
    oracle, and the root of the sum tree within 1e-7 relative after 20 write-backs."""
    mem, nmem, omem = ring
    st, mt = _trees(mem)
    omem.sum_tree.tree[:] = st
    omem.min_tree.tree[:] = mt
    eng, layout, oa, gen = _engine(13)
    worst = 0.0
    for step in range(20):
        u = torch.rand(B, generator=gen)
        z = (torch.randn(eng.noise_count, generator=gen), torch.randn(eng.noise_count, generator=gen))
        loss, idx, pri = _fused(eng, mem, nmem, u, z)
        oidx = omem.sample_proportional(B, uniforms=u)
        assert torch.equal(idx.cpu(), oidx), f"sampled indices drifted at step {step}"
        exp, nexp = _oracle_batch(nmem, oidx, omem.calculate_weights(oidx, BETA))
        oloss, _, opri = oa.learn_rainbow(exp, nexp, per=True, noise_normals=z)
        omem.update_priorities(oidx, opri)
        rel = abs(loss.item() - oloss) / max(1.0, abs(oloss))
        worst = max(worst, rel)
        assert rel <= (1e-5 if step == 0 else 1e-4), (step, loss.item(), oloss)
        np.testing.assert_allclose(pri.cpu().numpy(), np.asarray(opri).reshape(-1), rtol=1e-4, atol=1e-5)
    st1, _ = _trees(mem)
    assert abs(st1[1] - omem.sum_tree.tree[1]) <= 1e-7 * omem.sum_tree.tree[1]
    for k in oa.pkeys:                                   # 20 Adam steps of 1e-4: parameters still agree closely
        d = (eng.actor.view(k).cpu() - oa.actor[k].detach()).abs().max().item()
        assert d <= 2e-4, f"param {k} drifted by {d}"
    print(f"worst relative loss difference over 20 steps: {worst:.2e}")
