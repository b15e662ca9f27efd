"""Architecture descriptions + the flat HBM layout of one network.

``NetSpec`` says what the reference would build (``RainbowQNetwork`` q_networks.py:173-262,
``QNetwork`` :58-112 with encoders from networks/base.py:505-567); ``FlatLayout`` assigns every
tensor of the reference's ``state_dict()`` (same key names, evolvable_networks.py:577-643) a
slice of ONE flat fp32 parameter buffer (+ one flat epsilon buffer) and emits the
``b2rl_net_desc`` layer table the CUDA library consumes.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field

from .. import _lib


@dataclass
class MlpSpec:
    prefix: str
    name: str
    num_inputs: int
    num_outputs: int
    hidden_size: list
    noisy: bool = False
    layer_norm: bool = True
    output_layernorm: bool = False
    activation: str = "ReLU"
    output_activation: str | None = None


@dataclass
class CnnSpec:
    prefix: str
    name: str
    input_shape: tuple
    channel_size: list
    kernel_size: list
    stride_size: list
    num_outputs: int
    activation: str = "ReLU"
    output_activation: str | None = "ReLU"


@dataclass
class DenseSpec:
    """``EvolvableMultiInput`` over vector sub-spaces only (modules/multi_input.py:404-465): the raw vectors concatenated
    in key order -> ``final_dense`` -> output activation.  One linear layer under the reference's key."""
    key: str                        # "encoder.final_dense"
    num_inputs: int
    num_outputs: int
    output_activation: str | None = "ReLU"


@dataclass
class NetSpec:
    kind: str                       # "rainbow" | "q"
    encoder: object
    value: MlpSpec
    advantage: MlpSpec | None = None
    num_actions: int = 0
    num_atoms: int = 1
    obs_low: float | None = None
    obs_high: float | None = None
    obs_u8: bool = False


def conv_out(size: int, k: int, s: int) -> int:
    return (size - k) // s + 1


@dataclass
class Entry:
    key: str
    shape: tuple
    offset: int
    buf: str            # "param" | "eps"
    init: str = ""      # initialisation tag used by the modules


class FlatLayout:
    def __init__(self, spec: NetSpec):
        self.spec = spec
        self.entries: "OrderedDict[str, Entry]" = OrderedDict()
        self._np = 0                # buffer lengths (entries padded to 4 words; the padding stays zero)
        self._ne = 0
        self.n_param_elems = 0      # logical counts (what the reference's parameters()/buffers() hold)
        self.n_eps_elems = 0
        self.desc = _lib.NetDesc()
        d = self.desc
        d.kind = _lib.NET_RAINBOW if spec.kind == "rainbow" else _lib.NET_Q
        d.n_actions = spec.num_actions
        d.n_atoms = spec.num_atoms if spec.kind == "rainbow" else 1
        d.obs_u8 = int(spec.obs_u8)
        normalize = (isinstance(spec.encoder, CnnSpec) and spec.obs_low is not None
                     and not (spec.obs_low == 0.0 and spec.obs_high == 1.0))
        d.normalize = int(normalize)
        d.obs_low = float(spec.obs_low) if spec.obs_low is not None else 0.0
        d.obs_high = float(spec.obs_high) if spec.obs_high is not None else 1.0

        enc_layers = []
        if isinstance(spec.encoder, CnnSpec):
            e = spec.encoder
            c, h, w = e.input_shape
            d.obs_elems = c * h * w
            if len(e.channel_size) + 1 > _lib.B2RL_MAX_ENC:
                raise NotImplementedError("too many encoder layers for the CUDA layer table")
            for i, (co, k, s) in enumerate(zip(e.channel_size, e.kernel_size, e.stride_size), start=1):
                if isinstance(k, (tuple, list)):
                    if len(set(k)) != 1:
                        raise NotImplementedError("non-square conv kernels are not implemented in CUDA")
                    k = k[0]
                if isinstance(s, (tuple, list)):
                    s = s[0]
                oh, ow = conv_out(h, k, s), conv_out(w, k, s)
                if oh < 1 or ow < 1:
                    raise ValueError("convolution output collapsed to zero size")
                key = f"{e.prefix}{e.name}_conv_layer_{i}"
                L = self._layer(_lib.LAYER_CONV, e.activation, _lib.LN_NONE, False)
                L.in_c, L.in_h, L.in_w, L.out_c, L.out_h, L.out_w, L.ksize, L.stride = c, h, w, co, oh, ow, k, s
                L.w_off = self._param(key + ".weight", (co, c, k, k), "conv")
                L.b_off = self._param(key + ".bias", (co,), "zeros")
                enc_layers.append(L)
                c, h, w = co, oh, ow
            flat = c * h * w
            key = f"{e.prefix}{e.name}_linear_output"
            L = self._layer(_lib.LAYER_LINEAR, e.output_activation, _lib.LN_NONE, False)
            L.in_c, L.out_c = flat, e.num_outputs
            L.in_h = L.in_w = L.out_h = L.out_w = 1
            L.w_off = self._param(key + ".weight", (e.num_outputs, flat), "linear")
            L.b_off = self._param(key + ".bias", (e.num_outputs,), "linear_bias")
            enc_layers.append(L)
            self.flattened_size = flat
        elif isinstance(spec.encoder, DenseSpec):
            e = spec.encoder
            d.obs_elems = e.num_inputs
            enc_layers = [self._linear(e.key, e.num_inputs, e.num_outputs, e.output_activation, _lib.LN_NONE, False)]
        else:
            d.obs_elems = spec.encoder.num_inputs
            enc_layers = self._mlp_layers(spec.encoder)
        val_layers = self._mlp_layers(spec.value)
        adv_layers = self._mlp_layers(spec.advantage) if spec.advantage is not None else []
        if len(enc_layers) > _lib.B2RL_MAX_ENC or len(val_layers) > _lib.B2RL_MAX_HEAD or len(adv_layers) > _lib.B2RL_MAX_HEAD:
            raise NotImplementedError("network deeper than the CUDA layer table")
        d.n_enc, d.n_val, d.n_adv = len(enc_layers), len(val_layers), len(adv_layers)
        for i, L in enumerate(enc_layers):
            d.enc[i] = L
        for i, L in enumerate(val_layers):
            d.val[i] = L
        for i, L in enumerate(adv_layers):
            d.adv[i] = L
        d.n_params, d.n_eps = self._np, self._ne

    # -- helpers -------------------------------------------------------------------------------
    @staticmethod
    def _layer(kind, act, ln, noisy) -> _lib.Layer:
        if act not in _lib.ACT:
            raise NotImplementedError(f"activation {act!r} is not implemented in the CUDA kernels "
                                      f"(supported: {sorted(k for k in _lib.ACT if k)})")
        L = _lib.Layer()
        L.kind, L.act, L.ln, L.noisy = kind, _lib.ACT[act], ln, int(noisy)
        return L

    def _param(self, key, shape, init) -> int:
        n = 1
        for s in shape:
            n *= s
        off = self._np = (self._np + 3) & ~3      # 16-byte aligned views: float4 loads in the kernels
        self.entries[key] = Entry(key, tuple(shape), off, "param", init)
        self._np += n
        self.n_param_elems += n
        return off

    def _eps(self, key, shape) -> int:
        n = 1
        for s in shape:
            n *= s
        off = self._ne = (self._ne + 3) & ~3
        self.entries[key] = Entry(key, tuple(shape), off, "eps", "eps")
        self._ne += n
        self.n_eps_elems += n
        return off

    def _linear(self, key, n_in, n_out, act, ln, noisy, ln_key=None) -> _lib.Layer:
        L = self._layer(_lib.LAYER_LINEAR, act, ln, noisy)
        L.in_c, L.out_c = n_in, n_out
        L.in_h = L.in_w = L.out_h = L.out_w = 1
        if noisy:   # reference order: weight_mu, weight_sigma, bias_mu, bias_sigma, then buffers
            L.w_off = self._param(key + ".weight_mu", (n_out, n_in), "noisy_mu")
            L.ws_off = self._param(key + ".weight_sigma", (n_out, n_in), "noisy_wsigma")
            L.b_off = self._param(key + ".bias_mu", (n_out,), "noisy_mu_bias")
            L.bs_off = self._param(key + ".bias_sigma", (n_out,), "noisy_bsigma")
            L.we_off = self._eps(key + ".weight_epsilon", (n_out, n_in))
            L.be_off = self._eps(key + ".bias_epsilon", (n_out,))
        else:
            L.w_off = self._param(key + ".weight", (n_out, n_in), "linear")
            L.b_off = self._param(key + ".bias", (n_out,), "linear_bias")
        if ln == _lib.LN_AFFINE:
            L.lnw_off = self._param(ln_key + ".weight", (n_out,), "ones")
            L.lnb_off = self._param(ln_key + ".bias", (n_out,), "zeros")
        return L

    def _mlp_layers(self, m: MlpSpec) -> list:
        layers = []
        dims = [m.num_inputs, *m.hidden_size]
        for i in range(1, len(dims)):
            ln = _lib.LN_AFFINE if m.layer_norm else _lib.LN_NONE
            layers.append(self._linear(f"{m.prefix}{m.name}_linear_layer_{i}", dims[i - 1], dims[i], m.activation, ln,
                                       m.noisy, f"{m.prefix}{m.name}_layer_norm_{i}"))
        ln = _lib.LN_PLAIN if m.output_layernorm else _lib.LN_NONE
        layers.append(self._linear(f"{m.prefix}{m.name}_linear_layer_output", dims[-1], m.num_outputs,
                                   m.output_activation, ln, m.noisy))
        return layers

    @property
    def n_params(self) -> int:
        return self._np

    @property
    def n_eps(self) -> int:
        return self._ne

    def param_keys(self):
        return [k for k, e in self.entries.items() if e.buf == "param"]


def _prod(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


def rainbow_spec(obs_shape, num_actions, *, channel_size=(32, 32), kernel_size=(3, 3), stride_size=(1, 1),
                 latent_dim=32, hidden_size=(64,), num_atoms=51, activation="ReLU", head_activation=None,
                 obs_low=None, obs_high=None, obs_u8=False, encoder_hidden=(64, 64)) -> NetSpec:
    """What ``RainbowDQN`` builds (dqn_rainbow.py:191-218 -> q_networks.py:173-262): CNN encoder
    for image observations, plain LayerNorm MLP encoder otherwise; noisy dueling head with
    LayerNorm and no output activation."""
    head_activation = head_activation or activation
    if len(obs_shape) == 3:
        enc = CnnSpec("encoder.model.", "encoder", tuple(int(d) for d in obs_shape), list(channel_size),
                      list(kernel_size), list(stride_size), latent_dim, activation, activation)
    else:
        enc = MlpSpec("encoder.model.", "encoder", _prod(obs_shape), latent_dim, list(encoder_hidden), noisy=False,
                      layer_norm=True, output_layernorm=True, activation=activation, output_activation=activation)
        obs_low = obs_high = None
    val = MlpSpec("head_net.model.", "value", latent_dim, num_atoms, list(hidden_size), noisy=True,
                  layer_norm=True, activation=head_activation)
    adv = MlpSpec("head_net.advantage_net.", "advantage", latent_dim, num_actions * num_atoms, list(hidden_size),
                  noisy=True, layer_norm=True, activation=head_activation)
    return NetSpec("rainbow", enc, val, adv, num_actions, num_atoms, obs_low, obs_high, obs_u8)


def q_spec(obs_shape, num_actions, *, channel_size=(32, 32), kernel_size=(3, 3), stride_size=(1, 1), latent_dim=32,
           hidden_size=(32,), activation="ReLU", head_activation=None, obs_low=None, obs_high=None, obs_u8=False,
           encoder_hidden=(64, 64)) -> NetSpec:
    """What ``DQN`` builds (q_networks.py:58-112)."""
    head_activation = head_activation or activation
    if len(obs_shape) == 3:
        enc = CnnSpec("encoder.model.", "encoder", tuple(int(d) for d in obs_shape), list(channel_size),
                      list(kernel_size), list(stride_size), latent_dim, activation, activation)
    else:
        enc = MlpSpec("encoder.model.", "encoder", _prod(obs_shape), latent_dim, list(encoder_hidden), noisy=False,
                      layer_norm=True, output_layernorm=True, activation=activation, output_activation=activation)
        obs_low = obs_high = None
    val = MlpSpec("head_net.model.", "value", latent_dim, num_actions, list(hidden_size), noisy=False,
                  layer_norm=True, activation=head_activation)
    return NetSpec("q", enc, val, None, num_actions, 1, obs_low, obs_high, obs_u8)
