"""``EvolvableMLP`` — architecture description + mutations of agilerl/modules/mlp.py:10-336
(layer stack Linear/NoisyLinear -> LayerNorm -> activation built by
utils/evolvable_networks.py:527-644).  The contraction / LayerNorm / activation kernels are in
csrc/nn.cu; weights live in the owning network's flat buffers."""
from __future__ import annotations

from .base import EvolvableModule, MutationType, mutation


class EvolvableMLP(EvolvableModule):
    def __init__(self, num_inputs: int, num_outputs: int, hidden_size: list, activation: str = "ReLU",
                 output_activation: str | None = None, min_hidden_layers: int = 1, max_hidden_layers: int = 3,
                 min_mlp_nodes: int = 64, max_mlp_nodes: int = 500, layer_norm: bool = True,
                 output_layernorm: bool = False, output_vanish: bool = True, init_layers: bool = True,
                 noisy: bool = False, noise_std: float = 0.5, new_gelu: bool = False, device: str = "cuda",
                 name: str = "mlp", random_seed: int | None = None) -> None:
        super().__init__(device, random_seed)
        assert num_inputs > 0, "'num_inputs' cannot be less than or equal to zero."
        assert num_outputs > 0, "'num_outputs' cannot be less than or equal to zero."
        for n in hidden_size:
            assert n > 0, "'hidden_size' cannot contain zero, please enter a valid integer."
        assert len(hidden_size) != 0, "MLP must contain at least one hidden layer."
        assert min_hidden_layers < max_hidden_layers, "'min_hidden_layers' must be less than 'max_hidden_layers."
        assert min_mlp_nodes < max_mlp_nodes, "'min_mlp_nodes' must be less than 'max_mlp_nodes."
        if new_gelu:
            raise NotImplementedError("new_gelu is not implemented in the CUDA kernels")
        self.num_inputs, self.num_outputs = num_inputs, num_outputs
        self.hidden_size = list(hidden_size)
        self.activation, self.output_activation = activation, output_activation
        self.min_hidden_layers, self.max_hidden_layers = min_hidden_layers, max_hidden_layers
        self.min_mlp_nodes, self.max_mlp_nodes = min_mlp_nodes, max_mlp_nodes
        self.layer_norm, self.output_layernorm = layer_norm, output_layernorm
        self.output_vanish, self.init_layers = output_vanish, init_layers
        self.noisy, self.noise_std, self.new_gelu = noisy, noise_std, new_gelu
        self.name = name

    @mutation(MutationType.LAYER)
    def add_layer(self):
        """mlp.py:227-239."""
        if len(self.hidden_size) < self.max_hidden_layers:
            self.hidden_size += [self.hidden_size[-1]]
        else:
            return self.add_node()
        return None

    @mutation(MutationType.LAYER)
    def remove_layer(self):
        """mlp.py:241-252."""
        if len(self.hidden_size) > self.min_hidden_layers:
            self.hidden_size = self.hidden_size[:-1]
        else:
            return self.add_node()
        return None

    @mutation(MutationType.NODE)
    def add_node(self, hidden_layer: int | None = None, numb_new_nodes: int | None = None) -> dict:
        """mlp.py:254-282."""
        if hidden_layer is None:
            hidden_layer = int(self.rng.integers(0, len(self.hidden_size)))
        else:
            hidden_layer = min(hidden_layer, len(self.hidden_size) - 1)
        if numb_new_nodes is None:
            numb_new_nodes = int(self.rng.choice([16, 32, 64]))
        if self.hidden_size[hidden_layer] + numb_new_nodes <= self.max_mlp_nodes:
            self.hidden_size[hidden_layer] += numb_new_nodes
        return {"hidden_layer": hidden_layer, "numb_new_nodes": numb_new_nodes}

    @mutation(MutationType.NODE)
    def remove_node(self, hidden_layer: int | None = None, numb_new_nodes: int | None = None) -> dict:
        """mlp.py:284-312."""
        if hidden_layer is None:
            hidden_layer = int(self.rng.integers(0, len(self.hidden_size)))
        else:
            hidden_layer = min(hidden_layer, len(self.hidden_size) - 1)
        if numb_new_nodes is None:
            numb_new_nodes = int(self.rng.choice([16, 32, 64]))
        if self.hidden_size[hidden_layer] - numb_new_nodes > self.min_mlp_nodes:
            self.hidden_size[hidden_layer] -= numb_new_nodes
        return {"hidden_layer": hidden_layer, "numb_new_nodes": numb_new_nodes}
