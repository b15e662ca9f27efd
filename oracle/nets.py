"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement, in plain functional torch-fp32, of the reference's network forward for the
off-policy path.  Networks are described by a ``NetSpec`` and evaluated straight from a
state-dict keyed with the *reference's own parameter names*, so a reference ``state_dict()`` can
be fed in unchanged.

Follows (reference file:line):
  * conv stack + flatten + linear + activation   agilerl/utils/evolvable_networks.py:460-521,
                                                  agilerl/modules/cnn.py:487-546
  * MLP: (Noisy)Linear -> LayerNorm -> act ... -> output [-> LayerNorm(no affine)] -> out act
                                                  agilerl/utils/evolvable_networks.py:527-644
  * NoisyLinear  W = mu + sigma*eps (train mode)  agilerl/modules/custom_components.py:89-104
  * noise reset  eps_W = f(e_out) (x) f(e_in), f(x)=sign(x)sqrt|x|
                                                  agilerl/modules/custom_components.py:116-131
  * dueling distributional head                   agilerl/networks/custom_modules.py:127-162
  * image normalisation (x-low)/(high-low)        agilerl/utils/algo_utils.py:1131-1180
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


def activation(name: str | None, x: torch.Tensor) -> torch.Tensor:
    """agilerl/utils/evolvable_networks.py:348-374 (subset used by the mutation menu + Tanh)."""
    if name is None or name == "Identity":
        return x
    if name == "ReLU":
        return F.relu(x)
    if name == "ELU":
        return F.elu(x)
    if name == "GELU":
        return F.gelu(x)
    if name == "Tanh":
        return torch.tanh(x)
    raise NotImplementedError(name)


@dataclass
class MlpSpec:
    prefix: str                 # state-dict prefix, e.g. "head_net.model."
    name: str                   # layer-name stem, e.g. "value"
    num_inputs: int
    num_outputs: int
    hidden_size: list[int]
    noisy: bool = False
    layer_norm: bool = True
    output_layernorm: bool = False
    activation: str = "ReLU"
    output_activation: str | None = None


@dataclass
class CnnSpec:
    prefix: str                 # e.g. "encoder.model."
    name: str                   # e.g. "encoder"
    input_shape: tuple[int, int, int]
    channel_size: list[int]
    kernel_size: list[int]
    stride_size: list[int]
    num_outputs: int
    activation: str = "ReLU"
    output_activation: str | None = "ReLU"


@dataclass
class NetSpec:
    kind: str                               # "rainbow" | "q"
    encoder: CnnSpec | MlpSpec
    value: MlpSpec
    advantage: MlpSpec | None = None
    num_actions: int = 0
    num_atoms: int = 51
    # image normalisation (None -> no normalisation)
    obs_low: float | None = None
    obs_high: float | None = None
    support: torch.Tensor | None = field(default=None, repr=False)


def rainbow_spec(obs_shape, num_actions, channel_size=(32, 32), kernel_size=(8, 4),
                 stride_size=(4, 2), latent_dim=32, hidden_size=(64,), num_atoms=51,
                 activation="ReLU", obs_low=0.0, obs_high=255.0, encoder_hidden=(64, 64)) -> NetSpec:
    """Spec of RainbowQNetwork as RainbowDQN builds it (dqn_rainbow.py:191-218,
    q_networks.py:173-262): image obs -> EvolvableCNN encoder; vector obs -> EvolvableMLP encoder with
    layer_norm + output_layernorm (neither is noisy); noisy dueling head."""
    if len(obs_shape) == 3:
        enc = CnnSpec("encoder.model.", "encoder", tuple(obs_shape), list(channel_size),
                      list(kernel_size), list(stride_size), latent_dim, activation, activation)
    else:
        n_in = 1
        for d in obs_shape:
            n_in *= d
        # q_networks.py:189-206 sets noise_std/layer_norm on the encoder config but never
        # ``noisy``: the MLP encoder is a plain Linear stack (verified on the real reference)
        enc = MlpSpec("encoder.model.", "encoder", n_in, latent_dim, list(encoder_hidden), noisy=False,
                      layer_norm=True, output_layernorm=True, activation=activation,
                      output_activation=activation)
        obs_low = obs_high = None
    val = MlpSpec("head_net.model.", "value", latent_dim, num_atoms, list(hidden_size), noisy=True,
                  layer_norm=True, activation=activation)
    adv = MlpSpec("head_net.advantage_net.", "advantage", latent_dim, num_actions * num_atoms,
                  list(hidden_size), noisy=True, layer_norm=True, activation=activation)
    return NetSpec("rainbow", enc, val, adv, num_actions, num_atoms, obs_low, obs_high)


def q_spec(obs_shape, num_actions, channel_size=(32, 32), kernel_size=(3, 3), stride_size=(1, 1),
           latent_dim=32, hidden_size=(32,), activation="ReLU", obs_low=0.0, obs_high=255.0,
           encoder_hidden=(64, 64)) -> NetSpec:
    """Spec of QNetwork as DQN builds it (q_networks.py:58-112, networks/base.py:505-567)."""
    if len(obs_shape) == 3:
        enc = CnnSpec("encoder.model.", "encoder", tuple(obs_shape), list(channel_size),
                      list(kernel_size), list(stride_size), latent_dim, activation, activation)
    else:
        n_in = 1
        for d in obs_shape:
            n_in *= d
        enc = MlpSpec("encoder.model.", "encoder", n_in, latent_dim, list(encoder_hidden), noisy=False,
                      layer_norm=True, output_layernorm=True, activation=activation,
                      output_activation=activation)
        obs_low = obs_high = None
    val = MlpSpec("head_net.model.", "value", latent_dim, num_actions, list(hidden_size), noisy=False,
                  layer_norm=True, activation=activation)
    return NetSpec("q", enc, val, None, num_actions, 1, obs_low, obs_high)


# ---------------------------------------------------------------------------------------------
def _linear(sd, key: str, x: torch.Tensor, noisy: bool, train_noise: bool) -> torch.Tensor:
    if noisy:
        w, b = sd[key + ".weight_mu"], sd[key + ".bias_mu"]
        if train_noise:  # custom_components.py:97-99
            w = w + sd[key + ".weight_sigma"].mul(sd[key + ".weight_epsilon"])
            b = b + sd[key + ".bias_sigma"].mul(sd[key + ".bias_epsilon"])
        return F.linear(x, w, b)
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def mlp_forward(sd, spec: MlpSpec, x: torch.Tensor, train_noise: bool = True) -> torch.Tensor:
    """evolvable_networks.py:573-644 layer order."""
    p, n = spec.prefix, spec.name
    for i, h in enumerate(spec.hidden_size, start=1):
        x = _linear(sd, f"{p}{n}_linear_layer_{i}", x, spec.noisy, train_noise)
        if spec.layer_norm:
            x = F.layer_norm(x, (h,), sd[f"{p}{n}_layer_norm_{i}.weight"], sd[f"{p}{n}_layer_norm_{i}.bias"])
        x = activation(spec.activation, x)
    x = _linear(sd, f"{p}{n}_linear_layer_output", x, spec.noisy, train_noise)
    if spec.output_layernorm:
        x = F.layer_norm(x, (spec.num_outputs,))
    return activation(spec.output_activation, x)


def cnn_forward(sd, spec: CnnSpec, x: torch.Tensor) -> torch.Tensor:
    """cnn.py:487-546: conv->act ... flatten -> linear -> output act."""
    p, n = spec.prefix, spec.name
    for i, s in enumerate(spec.stride_size, start=1):
        x = F.conv2d(x, sd[f"{p}{n}_conv_layer_{i}.weight"], sd[f"{p}{n}_conv_layer_{i}.bias"], stride=s)
        x = activation(spec.activation, x)
    x = x.flatten(1)
    x = F.linear(x, sd[f"{p}{n}_linear_output.weight"], sd[f"{p}{n}_linear_output.bias"])
    return activation(spec.output_activation, x)


def preprocess(spec: NetSpec, obs: torch.Tensor) -> torch.Tensor:
    """core/base.py:1287-1301 -> algo_utils.py:993-1022: float(), image min-max, batch fix-up."""
    obs = obs.float()
    if isinstance(spec.encoder, CnnSpec):
        if spec.obs_low is not None and not (spec.obs_low == 0.0 and spec.obs_high == 1.0):
            low = torch.full(spec.encoder.input_shape, spec.obs_low, dtype=obs.dtype)
            high = torch.full(spec.encoder.input_shape, spec.obs_high, dtype=obs.dtype)
            obs = (obs - low) / (high - low)
        shape = spec.encoder.input_shape
    else:
        shape = (spec.encoder.num_inputs,)
    if obs.ndim == len(shape):
        obs = obs.unsqueeze(0)
    elif obs.ndim == len(shape) + 2:          # algo_utils.py:877-878 (quirk Q2 rescue)
        obs = obs.view(-1, *shape)
    return obs


def encode(sd, spec: NetSpec, obs: torch.Tensor, train_noise: bool = True) -> torch.Tensor:
    if isinstance(spec.encoder, CnnSpec):
        return cnn_forward(sd, spec.encoder, obs)
    return mlp_forward(sd, spec.encoder, obs.flatten(1), train_noise)


def rainbow_forward(sd, spec: NetSpec, obs: torch.Tensor, q: bool = True, log: bool = False,
                    train_noise: bool = True) -> torch.Tensor:
    """q_networks.py:265-284 + custom_modules.py:127-162 (obs already preprocessed)."""
    latent = encode(sd, spec, obs, train_noise)
    value = mlp_forward(sd, spec.value, latent, train_noise)
    adv = mlp_forward(sd, spec.advantage, latent, train_noise)
    B = value.size(0)
    value = value.view(B, 1, spec.num_atoms)
    adv = adv.view(B, spec.num_actions, spec.num_atoms)
    x = value + adv - adv.mean(1, keepdim=True)
    if log:
        x = F.log_softmax(x.view(-1, spec.num_atoms), dim=-1)
        return x.view(-1, spec.num_actions, spec.num_atoms)
    x = F.softmax(x.view(-1, spec.num_atoms), dim=-1)
    x = x.view(-1, spec.num_actions, spec.num_atoms).clamp(min=1e-3)
    if q:
        x = torch.sum(x * spec.support, dim=2)
    return x


def q_forward(sd, spec: NetSpec, obs: torch.Tensor) -> torch.Tensor:
    """q_networks.py:114-124."""
    return mlp_forward(sd, spec.value, encode(sd, spec, obs, False), False)


# ---------------------------------------------------------------------------------------------
def noisy_layer_keys(spec: NetSpec) -> list[tuple[str, int, int]]:
    """(state-dict key, in_features, out_features) of every NoisyLinear in module-traversal order
    (encoder, head value net, head advantage net) — the order ``reset_noise`` consumes RNG in
    (modules/base.py:573-577, networks/base.py reset_noise via children)."""
    out = []
    mlps = []
    if isinstance(spec.encoder, MlpSpec):
        mlps.append(spec.encoder)
    mlps.append(spec.value)
    if spec.advantage is not None:
        mlps.append(spec.advantage)
    for m in mlps:
        if not m.noisy:
            continue
        dims = [m.num_inputs, *m.hidden_size]
        for i in range(1, len(dims)):
            out.append((f"{m.prefix}{m.name}_linear_layer_{i}", dims[i - 1], dims[i]))
        out.append((f"{m.prefix}{m.name}_linear_layer_output", dims[-1], m.num_outputs))
    return out


def scale_noise(x: torch.Tensor) -> torch.Tensor:
    """custom_components.py:124-131."""
    return x.sign().mul_(x.abs().sqrt_())


def reset_noise_from_normals(sd, spec: NetSpec, normals: torch.Tensor) -> int:
    """Write eps buffers from a flat vector of standard normals laid out per layer as
    [randn(in), randn(out)] in traversal order (custom_components.py:116-122).  Returns the
    number of normals consumed."""
    off = 0
    for key, n_in, n_out in noisy_layer_keys(spec):
        e_in = scale_noise(normals[off:off + n_in].clone()); off += n_in
        e_out = scale_noise(normals[off:off + n_out].clone()); off += n_out
        sd[key + ".weight_epsilon"] = e_out.ger(e_in)
        sd[key + ".bias_epsilon"] = e_out.clone()
    return off


def num_noise_normals(spec: NetSpec) -> int:
    return sum(a + b for _, a, b in noisy_layer_keys(spec))


def param_keys(sd) -> list[str]:
    """Learnable keys = everything except the epsilon buffers."""
    return [k for k in sd if not k.endswith("_epsilon")]
