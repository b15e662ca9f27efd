"""Parameter initialisation with the reference's distributions, drawn from torch's CPU generator.

  conv            orthogonal(gain sqrt 2), bias 0          evolvable_networks.py:410-441 (layer_init)
  nn.Linear       torch default (kaiming-uniform a=sqrt 5) cnn.py:534-538, create_mlp init_layers=False
  NoisyLinear     mu ~ U(-1/sqrt(in), 1/sqrt(in)), sigma_W = std/sqrt(in), sigma_b = std/sqrt(out)
                                                            custom_components.py:106-114
  output_vanish   x0.1 on the output layer's mu / sigma    evolvable_networks.py:621-631
  LayerNorm       weight 1, bias 0
"""
from __future__ import annotations

import math

import torch


def _orthogonal_(t: torch.Tensor, gain: float) -> None:
    """``torch.nn.init.orthogonal_`` with the QR factorisation on ONE thread.  The matrices here are tiny ([32, 256] ..
    [64, 1024]); the LAPACK call behind ``torch.linalg.qr`` wakes the whole intra-op pool for them, which costs ~1 ms per
    layer on an idle many-core host and > 100 ms where the pool is oversubscribed (measured: 112 ms vs 0.04 ms for
    256 x 32) — per conv layer, per network, per ``clone()`` of every member of every generation.  Same draws from torch's
    generator (the ``normal_`` of the flattened matrix), same distribution; the values agree with a multi-threaded
    factorisation to ~2e-7 (the blocked and unblocked Householder variants round differently — as they already do between
    two machines running the reference)."""
    n = torch.get_num_threads()
    if n > 1:
        torch.set_num_threads(1)
    try:
        torch.nn.init.orthogonal_(t, gain)
    finally:
        if n > 1:
            torch.set_num_threads(n)


def init_state_dict(layout, noise_std: float = 0.5, output_vanish_heads: bool = True, init_mlp_layers: bool = False):
    sd = {}
    for key, e in layout.entries.items():
        if e.buf != "param":
            continue
        t = torch.empty(e.shape, dtype=torch.float32)
        if e.init == "conv":
            _orthogonal_(t, math.sqrt(2))
        elif e.init in ("zeros",):
            t.zero_()
        elif e.init == "ones":
            t.fill_(1.0)
        elif e.init == "linear":
            if init_mlp_layers:
                _orthogonal_(t, math.sqrt(2))
            else:
                torch.nn.init.kaiming_uniform_(t, a=math.sqrt(5))
        elif e.init == "linear_bias":
            if init_mlp_layers:
                t.zero_()
            else:
                wkey = key[: -len("bias")] + "weight"
                fan_in = layout.entries[wkey].shape[1]
                bound = 1.0 / math.sqrt(fan_in) if fan_in > 0 else 0.0
                t.uniform_(-bound, bound)
        elif e.init in ("noisy_mu", "noisy_wsigma"):
            fan_in = e.shape[1]
            if e.init == "noisy_mu":
                r = 1.0 / math.sqrt(fan_in)
                t.uniform_(-r, r)
            else:
                t.fill_(noise_std / math.sqrt(fan_in))
        elif e.init in ("noisy_mu_bias", "noisy_bsigma"):
            wkey = key.replace("bias_mu", "weight_mu").replace("bias_sigma", "weight_mu")
            fan_in = layout.entries[wkey].shape[1]
            if e.init == "noisy_mu_bias":
                r = 1.0 / math.sqrt(fan_in)
                t.uniform_(-r, r)
            else:
                t.fill_(noise_std / math.sqrt(e.shape[0]))
        else:
            raise AssertionError(e.init)
        if output_vanish_heads and "head_net" in key and "_linear_layer_output" in key:
            t.mul_(0.1)
        sd[key] = t
    return sd
