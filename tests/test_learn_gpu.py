"""GPU parity of the learn() hot path (forward x3, C51 projection, loss, backward, clip, Adam,
Polyak, noise reset) through the C ABI vs golden vectors from the unmodified reference and vs the
oracle.  Tolerance: losses / priorities / projected distributions within 1e-5 (north star)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

LOSS_TOL = dict(rtol=1e-5, atol=1e-5)


def _sd(g, tag):
    return {k[len(tag) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "/")}


def _engine(spec, g):
    from agilerl_b200.engine import LearnEngine, NetBuffers
    from agilerl_b200.networks.spec import FlatLayout
    layout = FlatLayout(spec)
    actor, target = NetBuffers(layout, "cuda"), NetBuffers(layout, "cuda")
    actor.load_state_dict(_sd(g, "actor0"))
    target.load_state_dict(_sd(g, "target0"))
    return LearnEngine(layout, actor, target), layout


def _batch(g, tag):
    return {k: torch.from_numpy(g[f"{tag}_{k}"]).cuda() for k in ("obs", "action", "reward", "next_obs", "done")}


def _adam_expected(p0, g, lr, step=1):
    m = 0.1 * g
    v = 0.001 * g * g
    bc1, bc2 = 1 - 0.9 ** step, 1 - 0.999 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + 1e-8
    return p0 - (lr / bc1) * m / denom


def _small_spec():
    from agilerl_b200.networks.spec import rainbow_spec
    return rainbow_spec((3, 20, 20), 4, channel_size=(8, 16), kernel_size=(4, 3), stride_size=(2, 1), latent_dim=16,
                        hidden_size=(32,), obs_low=0.0, obs_high=255.0, obs_u8=True)


def _run_rainbow(g, spec, *, n_step=True, combined=False, driver=False, weights_mode=1, with_state=True):
    B = int(g["B"])
    if with_state:
        eng, layout = _engine(spec, g)
    hp = dict(v_min=-10.0, v_max=10.0, delta_z=20.0 / 50, lr=1e-3, tau=1e-3, prior_eps=1e-6)
    support = torch.linspace(-10.0, 10.0, 51).cuda()
    passes = []
    if combined or not n_step:
        passes.append((_batch(g, "exp"), 0.99, False))
    if n_step:
        passes.append((_batch(g, "nexp"), 0.99 ** 3, driver))
    out = eng.rainbow_learn(passes, B=B, support=support, weights=torch.from_numpy(g["exp_weights"]).cuda(),
                            weights_mode=weights_mode, hp=hp,
                            noise_normals=(torch.from_numpy(g["z_actor"]), torch.from_numpy(g["z_target"])),
                            want_proj=True)
    torch.cuda.synchronize()
    return eng, layout, out


def test_forward_q_values_match_oracle():
    """get_action forward: expected Q-values with train-mode noise vs the oracle."""
    from oracle import nets as onets
    g = load_golden("rainbow_small_canonical.npz")
    eng, layout = _engine(_small_spec(), g)
    ospec = onets.rainbow_spec((3, 20, 20), 4, (8, 16), (4, 3), (2, 1), 16, (32,))
    ospec.support = torch.linspace(-10.0, 10.0, 51)
    obs = torch.from_numpy(g["exp_obs"])
    q_ref = onets.rainbow_forward(_sd(g, "actor0"), ospec, onets.preprocess(ospec, obs))
    q, am = eng.q_values(eng.actor, obs.cuda(), ospec.support.cuda(), use_noise=True, want_argmax=True)
    np.testing.assert_allclose(q.cpu().numpy(), q_ref.numpy(), rtol=1e-5, atol=1e-5)
    assert torch.equal(am.cpu(), q_ref.argmax(1))
    q_eval = eng.q_values(eng.actor, obs.cuda(), ospec.support.cuda(), use_noise=False)
    q_ref_eval = onets.rainbow_forward(_sd(g, "actor0"), ospec, onets.preprocess(ospec, obs), train_noise=False)
    np.testing.assert_allclose(q_eval.cpu().numpy(), q_ref_eval.numpy(), rtol=1e-5, atol=1e-5)


def test_forward_distributions_match_oracle():
    """RainbowQNetwork.forward(q=False) / (log=True): per-atom distributions and log-probabilities
    (custom_modules.py:153-160) vs the oracle, train-mode noise and eval mode."""
    from oracle import nets as onets
    g = load_golden("rainbow_small_canonical.npz")
    eng, layout = _engine(_small_spec(), g)
    ospec = onets.rainbow_spec((3, 20, 20), 4, (8, 16), (4, 3), (2, 1), 16, (32,))
    ospec.support = torch.linspace(-10.0, 10.0, 51)
    obs = torch.from_numpy(g["exp_obs"])
    x = onets.preprocess(ospec, obs)
    for noise in (True, False):
        ref = onets.rainbow_forward(_sd(g, "actor0"), ospec, x, q=False, train_noise=noise)
        got = eng.distributions(eng.actor, obs.cuda(), use_noise=noise)
        assert got.shape == (obs.shape[0], 4, 51)
        np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)
        assert float(got.min()) >= 1e-3                                   # clamp, quirk Q8
        ref_log = onets.rainbow_forward(_sd(g, "actor0"), ospec, x, q=False, log=True, train_noise=noise)
        got_log = eng.distributions(eng.actor, obs.cuda(), use_noise=noise, log=True)
        np.testing.assert_allclose(got_log.cpu().numpy(), ref_log.numpy(), rtol=1e-5, atol=1e-5)


def test_rainbow_learn_golden_canonical_full_state():
    g = load_golden("rainbow_small_canonical.npz")
    eng, layout, (loss, loss_elem, pri, proj) = _run_rainbow(g, _small_spec())
    np.testing.assert_allclose(proj.cpu().numpy(), g["proj_dist"], **LOSS_TOL)
    np.testing.assert_allclose(pri.cpu().numpy(), g["priorities"], **LOSS_TOL)
    np.testing.assert_allclose(loss.item(), float(g["loss"]), **LOSS_TOL)
    # gradients (after clip_grad_norm_) per tensor, names as in the reference state_dict
    gref = _sd(g, "grad")
    from agilerl_b200.engine import NetBuffers
    gv = NetBuffers(layout, "cuda"); gv.params.copy_(eng.grads)
    for k, ref in gref.items():
        got = gv.view(k).cpu()
        scale = max(ref.abs().max().item(), 1e-6)
        assert (got - ref).abs().max().item() <= 2e-5 * scale + 1e-8, f"grad {k}"
    # optimiser: Adam on OUR gradients must reproduce OUR parameters; Polyak exactly
    a0, t0, a1 = _sd(g, "actor0"), _sd(g, "target0"), _sd(g, "actor1")
    for k in gref:
        p1 = eng.actor.view(k).cpu()
        exp = _adam_expected(a0[k], gv.view(k).cpu(), 1e-3)
        np.testing.assert_allclose(p1.numpy(), exp.numpy(), rtol=0, atol=2e-7, err_msg=f"adam {k}")
        t1 = eng.target.view(k).cpu()
        np.testing.assert_array_equal(t1.numpy(), (1e-3 * p1 + (1.0 - 1e-3) * t0[k]).numpy())
        # and against the reference where the gradient is not vanishingly small
        mask = gref[k].abs() > 1e-5
        if mask.any():
            assert (p1 - a1[k])[mask].abs().max().item() <= 5e-6, f"param {k}"
        np.testing.assert_allclose(eng.exp_avg[layout.entries[k].offset:layout.entries[k].offset + ref_numel(gref[k])]
                                   .cpu().numpy().reshape(gref[k].shape), g[f"m/{k}"], rtol=1e-4, atol=1e-8)
    # noise reset from the injected normals == reference's new epsilon buffers.  Not bit-equal by
    # construction: torch's CPU sqrt goes through MKL VML, which is not correctly rounded
    # (0.6% of values are 1 ulp off IEEE sqrt); the device uses IEEE sqrt.rn.
    for k, v in a1.items():
        if k.endswith("_epsilon"):
            np.testing.assert_allclose(eng.actor.view(k).cpu().numpy(), v.numpy(), rtol=2.5e-7, atol=0)
            np.testing.assert_allclose(eng.target.view(k).cpu().numpy(), _sd(g, "target1")[k].numpy(), rtol=2.5e-7, atol=0)


def ref_numel(t):
    return t.numel()


@pytest.mark.parametrize("name,kw", [
    ("rainbow_small_wcol.npz", dict(weights_mode=2)),                       # quirk Q1: weights [B,1]
    ("rainbow_small_driver.npz", dict(weights_mode=2, driver=True)),        # quirks Q1+Q2: driver shapes
    ("rainbow_small_combined.npz", dict(combined=True)),
    ("rainbow_small_1step.npz", dict(n_step=False)),
])
def test_rainbow_learn_golden_variants(name, kw):
    """Needs initial weights: regenerate them exactly as make_golden.py did is not possible on the
    GPU box, so these cases carry outputs only and are checked CUDA-vs-oracle on the fixture's
    inputs with freshly initialised (seeded) weights shared by both sides."""
    from oracle import learn as olearn, nets as onets
    g = load_golden(name)
    B = int(g["B"])
    spec = _small_spec()
    ospec = onets.rainbow_spec((3, 20, 20), 4, (8, 16), (4, 3), (2, 1), 16, (32,))
    from agilerl_b200.engine import LearnEngine, NetBuffers
    from agilerl_b200.networks.spec import FlatLayout
    layout = FlatLayout(spec)
    gen = torch.Generator().manual_seed(123)
    sd_a, sd_t = {}, {}
    for k, e in layout.entries.items():
        scale = 0.3 if "epsilon" in k else (1.0 if "norm" in k and k.endswith("weight") else 0.08)
        base = torch.randn(e.shape, generator=gen) * scale + (1.0 if ("norm" in k and k.endswith("weight")) else 0.0)
        sd_a[k] = base
        sd_t[k] = base + 0.01 * torch.randn(e.shape, generator=gen)
    actor, target = NetBuffers(layout, "cuda"), NetBuffers(layout, "cuda")
    actor.load_state_dict(sd_a); target.load_state_dict(sd_t)
    eng = LearnEngine(layout, actor, target)
    hp = dict(v_min=-10.0, v_max=10.0, delta_z=20.0 / 50, lr=1e-3, tau=1e-3, prior_eps=1e-6)
    support = torch.linspace(-10.0, 10.0, 51)
    n_step, combined, driver = kw.get("n_step", True), kw.get("combined", False), kw.get("driver", False)
    wmode = kw.get("weights_mode", 1)
    exp = {k: torch.from_numpy(g[f"exp_{k}"]) for k in ("obs", "action", "reward", "next_obs", "done", "weights", "idxs")}
    nexp = {k: torch.from_numpy(g[f"nexp_{k}"]) for k in ("obs", "action", "reward", "next_obs", "done")}
    oa = olearn.OracleAgent(ospec, sd_a, sd_t, batch_size=B, lr=1e-3, combined_reward=combined)
    z = (torch.from_numpy(g["z_actor"]), torch.from_numpy(g["z_target"]))
    oloss, _, opri = oa.learn_rainbow(exp, nexp if n_step else None, per=True, noise_normals=z)
    passes = []
    if combined or not n_step:
        passes.append(({k: v.cuda() for k, v in exp.items()}, 0.99, False))
    if n_step:
        passes.append(({k: v.cuda() for k, v in nexp.items()}, 0.99 ** 3, driver))
    loss, loss_elem, pri, proj = eng.rainbow_learn(passes, B=B, support=support.cuda(), weights=exp["weights"].cuda(),
                                                   weights_mode=wmode, hp=hp, noise_normals=z, want_proj=True)
    scale = max(1.0, abs(oloss))
    assert abs(loss.item() - oloss) <= 1e-5 * scale, (loss.item(), oloss)
    np.testing.assert_allclose(pri.cpu().numpy(), opri, rtol=1e-5, atol=1e-5 * scale)
    np.testing.assert_allclose(proj.cpu().numpy(), oa.last_proj_dist.numpy(), rtol=1e-5, atol=1e-5 * scale)
    gv = NetBuffers(layout, "cuda"); gv.params.copy_(eng.grads)
    coef = 1.0
    for k in oa.pkeys:
        ref = oa.actor[k].grad
        got = gv.view(k).cpu()
        s = max(ref.abs().max().item(), 1e-6)
        assert (got - ref).abs().max().item() <= 5e-5 * s + 1e-8, f"grad {k}"


def test_rainbow_vector_obs_golden():
    from agilerl_b200.networks.spec import rainbow_spec
    g = load_golden("rainbow_vector.npz")
    spec = rainbow_spec((5,), 3, latent_dim=16, hidden_size=(32, 16), encoder_hidden=(24, 24))
    eng, layout, (loss, loss_elem, pri, proj) = _run_rainbow(g, spec)
    np.testing.assert_allclose(proj.cpu().numpy(), g["proj_dist"], **LOSS_TOL)
    np.testing.assert_allclose(pri.cpu().numpy(), g["priorities"], **LOSS_TOL)
    np.testing.assert_allclose(loss.item(), float(g["loss"]), **LOSS_TOL)
    gref = _sd(g, "grad")
    from agilerl_b200.engine import NetBuffers
    gv = NetBuffers(layout, "cuda"); gv.params.copy_(eng.grads)
    for k, ref in gref.items():
        s = max(ref.abs().max().item(), 1e-6)
        assert (gv.view(k).cpu() - ref).abs().max().item() <= 2e-5 * s + 1e-8, f"grad {k}"


@pytest.mark.parametrize("double", [0, 1])
def test_dqn_learn_golden(double):
    from agilerl_b200.networks.spec import q_spec
    g = load_golden(f"dqn_vector_double{double}.npz")
    B = int(g["B"])
    eng, layout = _engine(q_spec((4,), 2), g)
    exp = _batch(g, "exp")
    loss = eng.dqn_learn(exp, B=B, hp=dict(gamma=0.99, lr=1e-3, tau=1e-3), double=bool(double))
    np.testing.assert_allclose(loss.item(), float(g["loss"]), **LOSS_TOL)
    a1, t0 = _sd(g, "actor1"), _sd(g, "target0")
    a0 = _sd(g, "actor0")
    from agilerl_b200.engine import NetBuffers
    gv = NetBuffers(layout, "cuda"); gv.params.copy_(eng.grads)
    for k in a1:
        p1 = eng.actor.view(k).cpu()
        exp_p = _adam_expected(a0[k], gv.view(k).cpu(), 1e-3)
        np.testing.assert_allclose(p1.numpy(), exp_p.numpy(), rtol=0, atol=2e-7, err_msg=k)
        np.testing.assert_array_equal(eng.target.view(k).cpu().numpy(), (1e-3 * p1 + (1.0 - 1e-3) * t0[k]).numpy())
        step_ref = a1[k] - a0[k]
        mask = step_ref.abs() > 9e-4          # clear-sign Adam steps must agree with the reference
        if mask.any():
            assert (p1 - a1[k])[mask].abs().max().item() <= 2e-5, k


def test_northstar_b16_golden():
    """Reference architecture of BASELINE config 2 (4x84x84 uint8, conv 32/32 k8/4 s4/2, latent 32,
    head [64], 6 actions, 51 atoms) — outputs only."""
    from oracle import learn as olearn, nets as onets
    from agilerl_b200.networks.spec import rainbow_spec, FlatLayout
    from agilerl_b200.engine import LearnEngine, NetBuffers
    g = load_golden("rainbow_northstar_b16.npz")
    B = int(g["B"])
    spec = rainbow_spec((4, 84, 84), 6, channel_size=(32, 32), kernel_size=(8, 4), stride_size=(4, 2),
                        obs_low=0.0, obs_high=255.0, obs_u8=True)
    ospec = onets.rainbow_spec((4, 84, 84), 6)
    layout = FlatLayout(spec)
    assert layout.n_param_elems == 162730 and layout.n_eps_elems == 27429
    gen = torch.Generator().manual_seed(7)
    sd_a = {}
    for k, e in layout.entries.items():
        fan = e.shape[-1] if len(e.shape) > 1 else e.shape[0]
        if "norm" in k and k.endswith("weight"):
            sd_a[k] = torch.ones(e.shape)
        elif "epsilon" in k:
            sd_a[k] = torch.randn(e.shape, generator=gen) * 0.5
        elif "conv" in k and k.endswith("weight"):
            sd_a[k] = torch.randn(e.shape, generator=gen) * (1.0 / (e.shape[1] * e.shape[2] * e.shape[3]) ** 0.5)
        else:
            sd_a[k] = torch.randn(e.shape, generator=gen) * (0.5 / fan ** 0.5)
    sd_t = {k: v + 0.01 * torch.randn(v.shape, generator=gen) for k, v in sd_a.items()}
    actor, target = NetBuffers(layout, "cuda"), NetBuffers(layout, "cuda")
    actor.load_state_dict(sd_a); target.load_state_dict(sd_t)
    eng = LearnEngine(layout, actor, target)
    exp = {k: torch.from_numpy(g[f"exp_{k}"]) for k in ("obs", "action", "reward", "next_obs", "done", "weights", "idxs")}
    nexp = {k: torch.from_numpy(g[f"nexp_{k}"]) for k in ("obs", "action", "reward", "next_obs", "done")}
    z = (torch.from_numpy(g["z_actor"]), torch.from_numpy(g["z_target"]))
    oa = olearn.OracleAgent(ospec, sd_a, sd_t, batch_size=B, lr=1e-4)
    oloss, _, opri = oa.learn_rainbow(exp, nexp, per=True, noise_normals=z)
    hp = dict(v_min=-10.0, v_max=10.0, delta_z=20.0 / 50, lr=1e-4, tau=1e-3, prior_eps=1e-6)
    loss, _, pri, proj = eng.rainbow_learn([({k: v.cuda() for k, v in nexp.items()}, 0.99 ** 3, False)], B=B,
                                           support=torch.linspace(-10.0, 10.0, 51).cuda(), weights=exp["weights"].cuda(),
                                           weights_mode=1, hp=hp, noise_normals=z, want_proj=True)
    assert abs(loss.item() - oloss) <= 1e-5 * max(1.0, abs(oloss)), (loss.item(), oloss)
    np.testing.assert_allclose(pri.cpu().numpy(), opri, rtol=1e-5, atol=1e-5)
    gv = NetBuffers(layout, "cuda"); gv.params.copy_(eng.grads)
    for k in oa.pkeys:
        ref = oa.actor[k].grad
        s = max(ref.abs().max().item(), 1e-6)
        assert (gv.view(k).cpu() - ref).abs().max().item() <= 5e-5 * s + 1e-8, f"grad {k}"


def _random_state(layout, gen):
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


@pytest.mark.parametrize("name,obs,chan,kern,stride,u8", [
    ("k5s2_then_k3s1_odd", (3, 21, 21), (8, 12), (5, 3), (2, 1), True),         # k % s != 0, odd planes, T=3/S=1 dgrad
    ("three_layers_k5s2_dgrad", (2, 30, 30), (16, 24, 8), (4, 5, 3), (2, 2, 1), True),   # zero-padded parity taps
    ("wide_f32_cin40", (4, 16, 16), (40, 8), (3, 3), (1, 2), False),           # Cout 40 -> n_pad 48; k3 s2 dgrad over Cin 40
    ("k2s2_k1s1", (1, 12, 12), (4, 4), (2, 1), (2, 1), True),                  # 1x1 kernel, single-tap classes
])
def test_conv_geometry_sweep_matches_oracle(name, obs, chan, kern, stride, u8):
    """Forward, weight-gradient and input-gradient tensor-core kernels on geometries away from the
    north-star net (kernel not a multiple of the stride, odd planes, three conv layers, wide layers,
    float observations): loss / priorities within 1e-5, every gradient tensor within 2e-5 max|g|."""
    from oracle import learn as olearn, nets as onets
    from agilerl_b200.engine import LearnEngine, NetBuffers
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    B, A = 8, 3
    lo, hi = (0.0, 255.0) if u8 else (None, None)
    spec = rainbow_spec(obs, A, channel_size=chan, kernel_size=kern, stride_size=stride, latent_dim=16,
                        hidden_size=(24,), obs_low=lo, obs_high=hi, obs_u8=u8)
    ospec = onets.rainbow_spec(obs, A, channel_size=chan, kernel_size=kern, stride_size=stride, latent_dim=16,
                               hidden_size=(24,), obs_low=lo, obs_high=hi)
    layout = FlatLayout(spec)
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    sd_a = _random_state(layout, gen)
    sd_t = {k: v + 0.01 * torch.randn(v.shape, generator=gen) for k, v in sd_a.items()}
    actor, target = NetBuffers(layout, "cuda"), NetBuffers(layout, "cuda")
    actor.load_state_dict(sd_a); target.load_state_dict(sd_t)
    eng = LearnEngine(layout, actor, target)

    def frames():
        if u8:
            return torch.randint(0, 256, (B, *obs), dtype=torch.uint8, generator=gen)
        return torch.randn((B, *obs), generator=gen)
    nexp = dict(obs=frames(), action=torch.randint(0, A, (B,), generator=gen).float(),
                reward=torch.randn(B, 1, generator=gen), next_obs=frames(),
                done=(torch.rand(B, 1, generator=gen) < 0.2).float())
    exp = dict(nexp, weights=torch.rand(B, generator=gen) + 0.5, idxs=torch.arange(B))
    oa = olearn.OracleAgent(ospec, sd_a, sd_t, batch_size=B, lr=1e-4)
    z = (torch.randn(eng.noise_count, generator=gen), torch.randn(eng.noise_count, generator=gen))
    oloss, _, opri = oa.learn_rainbow(exp, nexp, per=True, noise_normals=z)
    hp = dict(v_min=-10.0, v_max=10.0, delta_z=20.0 / 50, lr=1e-4, tau=1e-3, prior_eps=1e-6)
    loss, _, pri, _ = eng.rainbow_learn([({k: v.cuda() for k, v in nexp.items()}, 0.99 ** 3, False)], B=B,
                                        support=torch.linspace(-10.0, 10.0, 51).cuda(), weights=exp["weights"].cuda(),
                                        weights_mode=1, hp=hp, noise_normals=z)
    torch.cuda.synchronize()
    assert abs(loss.item() - oloss) <= 1e-5 * max(1.0, abs(oloss)), (loss.item(), oloss)
    np.testing.assert_allclose(pri.cpu().numpy(), np.asarray(opri).reshape(-1), rtol=1e-5, atol=1e-5)
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in oa.last_grads.values())).item()
    assert total < 10.0, "test weights must stay below the clip threshold"
    gv = NetBuffers(layout, "cuda"); gv.params.copy_(eng.grads)
    for k, ref in oa.last_grads.items():
        got = gv.view(k).cpu()
        scale = max(ref.abs().max().item(), 1e-6)
        assert (got - ref).abs().max().item() <= 2e-5 * scale + 1e-8, f"grad {k}: {(got - ref).abs().max().item()} vs {scale}"


@pytest.mark.parametrize("B", [1, 7, 300])
def test_odd_batch_sizes_match_oracle(B):
    """Batch sizes that are not multiples of the head tile (4 rows), the warp or the conv M tile:
    a single row, a prime, and one above a tile boundary — loss / priorities / gradients vs the oracle."""
    from oracle import learn as olearn, nets as onets
    from agilerl_b200.engine import LearnEngine, NetBuffers
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    obs, A = (3, 20, 20), 4
    kw = dict(channel_size=(8, 16), kernel_size=(4, 3), stride_size=(2, 1), latent_dim=16, hidden_size=(32,),
              obs_low=0.0, obs_high=255.0)
    layout = FlatLayout(rainbow_spec(obs, A, obs_u8=True, **kw))
    gen = torch.Generator().manual_seed(100 + B)
    sd_a = _random_state(layout, gen)
    sd_t = {k: v + 0.01 * torch.randn(v.shape, generator=gen) for k, v in sd_a.items()}
    actor, target = NetBuffers(layout, "cuda"), NetBuffers(layout, "cuda")
    actor.load_state_dict(sd_a); target.load_state_dict(sd_t)
    eng = LearnEngine(layout, actor, target)
    nexp = dict(obs=torch.randint(0, 256, (B, *obs), dtype=torch.uint8, generator=gen),
                action=torch.randint(0, A, (B,), generator=gen).float(), reward=torch.randn(B, 1, generator=gen),
                next_obs=torch.randint(0, 256, (B, *obs), dtype=torch.uint8, generator=gen),
                done=(torch.rand(B, 1, generator=gen) < 0.2).float())
    exp = dict(nexp, weights=torch.rand(B, generator=gen) + 0.5, idxs=torch.arange(B))
    oa = olearn.OracleAgent(onets.rainbow_spec(obs, A, **kw), sd_a, sd_t, batch_size=B, lr=1e-4)
    z = (torch.randn(eng.noise_count, generator=gen), torch.randn(eng.noise_count, generator=gen))
    oloss, _, opri = oa.learn_rainbow(exp, nexp, per=True, noise_normals=z)
    hp = dict(v_min=-10.0, v_max=10.0, delta_z=20.0 / 50, lr=1e-4, tau=1e-3, prior_eps=1e-6)
    loss, _, pri, _ = eng.rainbow_learn([({k: v.cuda() for k, v in nexp.items()}, 0.99 ** 3, False)], B=B,
                                        support=torch.linspace(-10.0, 10.0, 51).cuda(), weights=exp["weights"].cuda(),
                                        weights_mode=1, hp=hp, noise_normals=z)
    torch.cuda.synchronize()
    assert abs(loss.item() - oloss) <= 1e-5 * max(1.0, abs(oloss)), (loss.item(), oloss)
    np.testing.assert_allclose(pri.cpu().numpy(), np.asarray(opri).reshape(-1), rtol=1e-5, atol=1e-5)
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in oa.last_grads.values())).item()
    coef = min(1.0, 10.0 / (total + 1e-6))                   # clip_grad_norm_ scales the stored gradients
    gv = NetBuffers(layout, "cuda"); gv.params.copy_(eng.grads)
    for k, ref in oa.last_grads.items():
        got = gv.view(k).cpu()
        ref = ref * coef
        scale = max(ref.abs().max().item(), 1e-6)
        assert (got - ref).abs().max().item() <= 2e-5 * scale + 1e-8, f"grad {k}"


@pytest.mark.parametrize("cin,hw,k,s,cout,rows", [
    (4, 84, 8, 4, 32, 37),       # the north-star first layer; 37 images: ragged last tile, 116 tiles < 148 SMs
    (4, 84, 8, 4, 32, 300),      # 938 tiles: every CTA of the persistent kernel walks 6-7 tiles through both stages
    (2, 36, 4, 4, 16, 5),        # k4: one 4-byte group per kernel row, a single k-step, 16 channels
    (4, 40, 8, 4, 24, 9),        # Cout not a multiple of 16 (padded digit planes)
    (4, 64, 8, 8, 64, 3),        # 64 channels: 256 accumulator columns for one weight set
    (8, 36, 8, 4, 16, 6),        # K = 512: float digit recombination (the integer pairing needs K <= 256)
    (3, 40, 8, 4, 24, 4),        # odd channel count: outside the integer path, falls back to the tf32 kernel
])
def test_first_layer_int8_digit_conv_matches_float64(cin, hw, k, s, cout, rows):
    """conv_fwd_i8_kernel (tcgen05 kind::i8 over raw frame bytes, weights as four int8 digit planes) through the
    layer hook, against a float64 convolution of (x - low) / (high - low): <= 2e-6 of the output scale — tighter
    than the fp32 reference itself — for gathered ring rows, a ragged last tile and padded channel counts."""
    import ctypes
    from agilerl_b200 import _lib
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    spec = rainbow_spec((cin, hw, hw), 3, channel_size=(cout,), kernel_size=(k,), stride_size=(s,), latent_dim=16,
                        hidden_size=(16,), obs_low=0.0, obs_high=255.0, obs_u8=True)
    layout = FlatLayout(spec)
    desc = layout.desc
    L = desc.enc[0]
    params = torch.zeros(layout.n_params)
    w = torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k) ** 0.5)
    w[0] *= 1e-3                                      # a channel far below the others: per-channel scales matter
    w[1, :, :, : k // 2] = 0.0
    b = torch.randn(cout, generator=g) * 0.1
    params[L.w_off:L.w_off + w.numel()] = w.reshape(-1)
    params[L.b_off:L.b_off + cout] = b
    ring = torch.randint(0, 256, (64, cin, hw, hw), dtype=torch.uint8, generator=g)
    idx = torch.randint(0, 64, (rows,), generator=g)
    ref = torch.nn.functional.conv2d(ring[idx].double() / 255.0, w.double(), b.double(), stride=s).relu()
    out = torch.empty(ref.shape, dtype=torch.float32, device="cuda")
    ws = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
    pd, rd, id_ = params.cuda(), ring.cuda(), idx.cuda()
    lib = _lib.load()
    n0 = lib.b2rl_launch_count()
    _lib.check(lib.b2rl_encoder_layer_forward(ctypes.byref(desc), 0, pd.data_ptr(), rd.data_ptr(), id_.data_ptr(), rows,
                                              out.data_ptr(), ws.data_ptr(), ws.numel(), 0, _lib.stream_ptr(torch.device("cuda:0"))))
    torch.cuda.synchronize()
    assert lib.b2rl_launch_count() - n0 == 2, "expected the digit-split + int8 convolution launches"
    err = (out.cpu().double() - ref).abs().max().item()
    assert err <= 2e-6 * max(1.0, ref.abs().max().item()), err
    # low-magnitude channel: relative accuracy is kept by the per-channel scale
    assert (out.cpu().double()[:, 0] - ref[:, 0]).abs().max().item() <= 2e-6 * max(1e-3, ref[:, 0].abs().max().item())


@pytest.mark.parametrize("obs_hw,c0,k0,s0,cout,k,s,rows,staged", [
    (84, 32, 8, 4, 32, 4, 2, 256, True),     # the north-star second layer at the benchmark batch: 296 tiles of 71 pixels
    (84, 32, 8, 4, 32, 4, 2, 512, True),     # online pass over [next_obs | obs]: four tiles per CTA, slab refilled per channel
    (84, 32, 8, 4, 32, 4, 2, 1, True),       # one image: fewer tiles than CTAs
    (84, 32, 8, 4, 32, 4, 2, 37, True),      # ragged last tile
    (34, 8, 4, 2, 16, 8, 2, 9, True),        # k8: four k-blocks per channel, two 4-tap chunks per kernel row
    (50, 16, 4, 2, 48, 4, 4, 5, True),       # stride 4 (16-byte aligned taps), Cout 48
    (26, 2, 4, 2, 8, 4, 2, 300, True),       # K = 32: two k-blocks, one weight group; tiles span several 5x5 images
    (44, 4, 4, 2, 8, 4, 2, 3, False),        # input width 21: rows not 16-byte multiples -> gather kernel
])
def test_inner_conv_layer_staged_matches_float64(obs_hw, c0, k0, s0, cout, k, s, rows, staged):
    """conv_fwd_st_kernel (receptive fields bulk-copied into shared memory, 3xTF32 on tcgen05) through the layer hook,
    against a float64 convolution: <= 1e-5 of the output scale (tcgen05 accumulates in fp32 with truncation, K/8 k-steps of
    three products each: a bias of up to ~2^-24 of the running sum per accumulation, the same for the gather kernel);
    the path counter proves which kernel ran."""
    import ctypes
    from agilerl_b200 import _lib
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    g = torch.Generator().manual_seed(obs_hw * 100 + cout + rows)
    spec = rainbow_spec((4, obs_hw, obs_hw), 3, channel_size=(c0, cout), kernel_size=(k0, k), stride_size=(s0, s),
                        latent_dim=16, hidden_size=(16,), obs_low=0.0, obs_high=255.0, obs_u8=True)
    layout = FlatLayout(spec)
    desc = layout.desc
    L = desc.enc[1]
    hw = (obs_hw - k0) // s0 + 1
    assert (L.in_c, L.in_h, L.in_w) == (c0, hw, hw)
    params = torch.zeros(layout.n_params)
    w = torch.randn(cout, c0, k, k, generator=g) * (1.0 / (c0 * k * k) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    params[L.w_off:L.w_off + w.numel()] = w.reshape(-1)
    params[L.b_off:L.b_off + cout] = b
    x = torch.randn(rows, c0, hw, hw, generator=g).relu()            # post-ReLU activations of the layer before
    x[0, 0, 0, :4] = torch.tensor([1e-3, 1.0, 30.0, 3.0])             # mixed magnitudes inside one 4-tap chunk
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s).relu()
    out = torch.full(ref.shape, float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
    pd, xd = params.cuda(), x.cuda()
    lib = _lib.load()
    n_st, n_tc = lib.b2rl_conv_path_count(2), lib.b2rl_conv_path_count(0)
    prev = lib.b2rl_conv_staged_paths(7)
    try:
        _lib.check(lib.b2rl_encoder_layer_forward(ctypes.byref(desc), 1, pd.data_ptr(), xd.data_ptr(), None, rows,
                                                  out.data_ptr(), ws.data_ptr(), ws.numel(), 0, _lib.stream_ptr(torch.device("cuda:0"))))
        torch.cuda.synchronize()
    finally:
        lib.b2rl_conv_staged_paths(prev)
    assert lib.b2rl_conv_path_count(2) - n_st == (1 if staged else 0)
    assert lib.b2rl_conv_path_count(0) - n_tc == (0 if staged else 1)
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= 1e-5 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("first,obs_hw,c0,k0,s0,cout,k,s,rows,staged", [
    (False, 84, 32, 8, 4, 32, 4, 2, 256, True),   # north-star second layer: K = 512 taps (4 tap tiles), 81 pixels per image
    (False, 84, 32, 8, 4, 32, 4, 2, 37, True),    # ragged last tile, fewer tiles than CTAs
    (True, 84, 32, 8, 4, 32, 4, 2, 256, True),    # north-star first layer: uint8 frames of gathered ring rows, exact operand
    (True, 84, 32, 8, 4, 32, 4, 2, 5, True),
    (False, 34, 8, 4, 2, 16, 8, 2, 9, True),      # k8 over fp32, Cout 16
    (False, 50, 16, 4, 2, 48, 4, 4, 5, False),    # Cout 48: 128 G threads do not divide -> gather kernel
    (False, 26, 2, 4, 2, 8, 4, 2, 300, True),     # K = 32: a quarter of one tap tile
    (True, 36, 16, 4, 4, 8, 4, 2, 6, True),       # first layer k4 s4: K = 64
])
def test_conv_weight_gradient_staged_matches_float64(first, obs_hw, c0, k0, s0, cout, k, s, rows, staged):
    """The layer's dW / db alone against float64 autograd: conv_wgrad_st_kernel (staged operands, transposed im2col reads) for
    fp32 activations, conv_wgrad_i8_kernel (frame bytes as the MN-major operand, G as five int8 digit planes, int64
    accumulation) for the first layer's uint8 frames."""
    import ctypes
    from agilerl_b200 import _lib
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    g = torch.Generator().manual_seed(obs_hw * 7 + cout + rows + int(first))
    spec = rainbow_spec((4, obs_hw, obs_hw), 3, channel_size=(c0, cout), kernel_size=(k0, k), stride_size=(s0, s),
                        latent_dim=16, hidden_size=(16,), obs_low=0.0, obs_high=255.0, obs_u8=True)
    layout = FlatLayout(spec)
    desc = layout.desc
    li = 0 if first else 1
    L = desc.enc[li]
    if first:
        ring = torch.randint(0, 256, (64, L.in_c, L.in_h, L.in_w), dtype=torch.uint8, generator=g)
        idx = torch.randint(0, 64, (rows,), generator=g)
        x64 = ring[idx].double() / 255.0
        x_dev, idx_dev = ring.cuda(), idx.cuda()
    else:
        x = torch.randn(rows, L.in_c, L.in_h, L.in_w, generator=g).relu()
        x64 = x.double()
        x_dev, idx_dev = x.cuda(), None
    gout = torch.randn(rows, L.out_c, L.out_h, L.out_w, generator=g) * 0.1
    gout[:, 0] *= 1e-3
    w64 = torch.zeros(L.out_c, L.in_c, L.ksize, L.ksize, dtype=torch.float64, requires_grad=True)
    b64 = torch.zeros(L.out_c, dtype=torch.float64, requires_grad=True)
    out = torch.nn.functional.conv2d(x64, w64, b64, stride=L.stride)
    out.backward(gout.double())
    grads = torch.full((layout.n_params,), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    gdev = gout.cuda()
    lib = _lib.load()
    n_st, n_i8 = lib.b2rl_conv_path_count(2), lib.b2rl_conv_path_count(1)
    prev = lib.b2rl_conv_staged_paths(7)
    try:
        _lib.check(lib.b2rl_encoder_layer_wgrad(ctypes.byref(desc), li, x_dev.data_ptr(),
                                                idx_dev.data_ptr() if idx_dev is not None else None, rows, gdev.data_ptr(),
                                                grads.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _lib.stream_ptr(torch.device("cuda:0"))))
        torch.cuda.synchronize()
    finally:
        lib.b2rl_conv_staged_paths(prev)
    if first:      # uint8 frames: the integer-tensor weight gradient (conv_wi8.cuh) takes the layer
        assert lib.b2rl_conv_path_count(1) - n_i8 == 1 and lib.b2rl_conv_path_count(2) == n_st
    else:
        assert lib.b2rl_conv_path_count(2) - n_st == (1 if staged else 0)
    gh = grads.cpu().double()
    got_w, got_b = gh[L.w_off:L.w_off + w64.numel()].reshape(w64.shape), gh[L.b_off:L.b_off + L.out_c]
    assert torch.isfinite(got_w).all() and torch.isfinite(got_b).all()
    sw, sb = max(w64.grad.abs().max().item(), 1e-6), max(b64.grad.abs().max().item(), 1e-6)
    assert (got_w - w64.grad).abs().max().item() <= 5e-6 * sw, ((got_w - w64.grad).abs().max().item(), sw)
    assert (got_b - b64.grad).abs().max().item() <= 5e-6 * sb
    # the low-magnitude channel keeps its relative accuracy
    s0_ = max(w64.grad[0].abs().max().item(), 1e-9)
    assert (got_w[0] - w64.grad[0]).abs().max().item() <= 2e-5 * s0_


@pytest.mark.parametrize("obs_hw,c0,k0,s0,cout,k,s,rows,staged", [
    (84, 32, 8, 4, 32, 4, 2, 256, True),     # north-star second layer: 25 600 positions x (4 classes x 32 channels)
    (84, 32, 8, 4, 32, 4, 2, 3, True),       # fewer tiles than CTAs
    (84, 32, 8, 4, 32, 4, 2, 77, True),      # ragged
    (50, 16, 4, 2, 8, 4, 2, 5, True),        # 24x24 input, 16 input channels (64 columns), 8 output channels (K = 32)
    (42, 8, 4, 2, 16, 4, 2, 6, True),        # 20x20 -> 9x9, Cin 8 padded to 16
    (34, 8, 4, 2, 16, 8, 2, 9, False),       # k8 s2: T = 4, per-class kernel
])
def test_conv_input_gradient_staged_matches_float64(obs_hw, c0, k0, s0, cout, k, s, rows, staged):
    """conv_dgrad_st_kernel (the four parity classes as extra columns of one product) against float64 autograd."""
    import ctypes
    from agilerl_b200 import _lib
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    g = torch.Generator().manual_seed(obs_hw * 3 + cout + rows)
    spec = rainbow_spec((4, obs_hw, obs_hw), 3, channel_size=(c0, cout), kernel_size=(k0, k), stride_size=(s0, s),
                        latent_dim=16, hidden_size=(16,), obs_low=0.0, obs_high=255.0, obs_u8=True)
    layout = FlatLayout(spec)
    desc = layout.desc
    L = desc.enc[1]
    params = torch.zeros(layout.n_params)
    w = torch.randn(cout, c0, k, k, generator=g) * (1.0 / (cout * k * k) ** 0.5)
    params[L.w_off:L.w_off + w.numel()] = w.reshape(-1)
    gout = torch.randn(rows, L.out_c, L.out_h, L.out_w, generator=g)
    x64 = torch.zeros(rows, L.in_c, L.in_h, L.in_w, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(x64, w.double(), None, stride=s).backward(gout.double())
    gin = torch.full(x64.shape, float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
    pd, gd = params.cuda(), gout.cuda()
    lib = _lib.load()
    n_st = lib.b2rl_conv_path_count(2)
    prev = lib.b2rl_conv_staged_paths(7)
    try:
        _lib.check(lib.b2rl_encoder_layer_dgrad(ctypes.byref(desc), 1, pd.data_ptr(), gd.data_ptr(), rows, gin.data_ptr(),
                                                ws.data_ptr(), ws.numel(), _lib.stream_ptr(torch.device("cuda:0"))))
        torch.cuda.synchronize()
    finally:
        lib.b2rl_conv_staged_paths(prev)
    assert lib.b2rl_conv_path_count(2) - n_st == (1 if staged else 0)
    got = gin.cpu().double()
    assert torch.isfinite(got).all()
    err = (got - x64.grad).abs().max().item()
    assert err <= 5e-6 * max(1.0, x64.grad.abs().max().item()), err


def test_first_layer_weight_gradient_low_bound_zero_and_tiny_channels():
    """The integer weight gradient with a non-zero integer lower bound (the -low * sum(G) correction), an all-zero
    gradient channel and a channel six orders of magnitude below the others (per-channel scales)."""
    import ctypes
    from agilerl_b200 import _lib
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    g = torch.Generator().manual_seed(7)
    spec = rainbow_spec((4, 84, 84), 3, channel_size=(32, 32), kernel_size=(8, 4), stride_size=(4, 2), latent_dim=16,
                        hidden_size=(16,), obs_low=-3.0, obs_high=252.0, obs_u8=True)
    layout = FlatLayout(spec)
    desc = layout.desc
    L = desc.enc[0]
    rows = 33
    ring = torch.randint(0, 256, (48, L.in_c, L.in_h, L.in_w), dtype=torch.uint8, generator=g)
    idx = torch.randint(0, 48, (rows,), generator=g)
    x64 = (ring[idx].double() + 3.0) / 255.0
    gout = torch.randn(rows, L.out_c, L.out_h, L.out_w, generator=g) * 0.05
    gout[:, 3] = 0.0                                     # an all-zero channel
    gout[:, 5] *= 1e-6
    w64 = torch.zeros(L.out_c, L.in_c, L.ksize, L.ksize, dtype=torch.float64, requires_grad=True)
    b64 = torch.zeros(L.out_c, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.conv2d(x64, w64, b64, stride=L.stride).backward(gout.double())
    lib = _lib.load()
    grads = torch.full((layout.n_params,), float("nan"), dtype=torch.float32, device="cuda")
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    n_i8 = lib.b2rl_conv_path_count(1)
    ring_d, idx_d, gout_d = ring.cuda(), idx.cuda(), gout.cuda()
    _lib.check(lib.b2rl_encoder_layer_wgrad(ctypes.byref(desc), 0, ring_d.data_ptr(), idx_d.data_ptr(), rows,
                                            gout_d.data_ptr(), grads.data_ptr(), ws.data_ptr(), ws.numel(),
                                            _lib.stream_ptr(torch.device("cuda:0"))))
    torch.cuda.synchronize()
    assert lib.b2rl_conv_path_count(1) - n_i8 == 1
    gh = grads.cpu().double()
    got_w, got_b = gh[L.w_off:L.w_off + w64.numel()].reshape(w64.shape), gh[L.b_off:L.b_off + L.out_c]
    sw = w64.grad.abs().max().item()
    assert (got_w - w64.grad).abs().max().item() <= 2e-6 * sw
    assert (got_b - b64.grad).abs().max().item() <= 2e-6 * b64.grad.abs().max().item()
    assert got_w[3].abs().max().item() == 0.0 and got_b[3].item() == 0.0
    s5 = w64.grad[5].abs().max().item()
    assert (got_w[5] - w64.grad[5]).abs().max().item() <= 2e-6 * s5          # per-channel scale keeps small channels exact
