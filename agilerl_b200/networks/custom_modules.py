"""``DuelingDistributionalMLP`` — description of the Rainbow head
(agilerl/networks/custom_modules.py:11-188): a value MLP (name "value", ``num_atoms`` outputs) and
an advantage MLP (name "advantage", ``num_outputs*num_atoms`` outputs) with identical hidden
sizes; mutations applied to one apply to both.  The dueling combine / softmax / clamp /
expectation kernels are ``rainbow_q_kernel`` & co. in csrc/nn.cu."""
from __future__ import annotations

from ..modules.mlp import EvolvableMLP


class DuelingDistributionalMLP(EvolvableMLP):
    def __init__(self, num_inputs: int, num_outputs: int, hidden_size: list, num_atoms: int, support,
                 layer_norm: bool = True, output_layernorm: bool = False, output_vanish: bool = True,
                 init_layers: bool = False, noisy: bool = True, noise_std: float = 0.5, activation: str = "ReLU",
                 output_activation: str | None = None, min_hidden_layers: int = 1, max_hidden_layers: int = 3,
                 min_mlp_nodes: int = 64, max_mlp_nodes: int = 500, new_gelu: bool = False,
                 device: str = "cuda", random_seed: int | None = None) -> None:
        super().__init__(num_inputs, num_atoms, hidden_size, activation, output_activation, min_hidden_layers,
                         max_hidden_layers, min_mlp_nodes, max_mlp_nodes, layer_norm=layer_norm,
                         output_layernorm=output_layernorm, output_vanish=output_vanish, init_layers=init_layers,
                         noisy=noisy, noise_std=noise_std, new_gelu=new_gelu, device=device, name="value",
                         random_seed=random_seed)
        self.num_atoms = num_atoms
        self.num_actions = num_outputs
        self.support = support

    @property
    def init_dict(self):
        d = super().init_dict
        d["num_outputs"] = self.num_actions       # the ctor's num_outputs is the action count
        d.pop("name", None)
        return d

    @property
    def net_config(self):
        cfg = super().net_config
        cfg.pop("num_atoms", None)
        cfg.pop("support", None)
        return cfg
