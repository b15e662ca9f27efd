"""``EvolvableCNN`` — architecture description + mutations of agilerl/modules/cnn.py:224-788
(Conv2d -> activation ... -> Flatten -> Linear -> output activation, built by
utils/evolvable_networks.py:460-521).  Only Conv2d without BatchNorm is implemented in CUDA
(the configuration RainbowDQN / DQN build for image observations)."""
from __future__ import annotations

import numpy as np

from .base import EvolvableModule, MutationType, mutation


def _conv_out(size: int, k: int, s: int) -> int:
    return (size - k) // s + 1


class EvolvableCNN(EvolvableModule):
    def __init__(self, input_shape, num_outputs: int, channel_size: list, kernel_size: list, stride_size: list,
                 sample_input=None, block_type: str = "Conv2d", activation: str = "ReLU",
                 output_activation: str | None = None, min_hidden_layers: int = 1, max_hidden_layers: int = 6,
                 min_channel_size: int = 16, max_channel_size: int = 256, layer_norm: bool = False,
                 init_layers: bool = True, device: str = "cuda", name: str = "cnn",
                 random_seed: int | None = None) -> None:
        super().__init__(device, random_seed)
        assert len(kernel_size) == len(channel_size), (
            "Length of kernel size list must be the same length as channel size list.")
        assert len(stride_size) == len(channel_size), (
            "Length of stride size list must be the same length as channel size list.")
        assert num_outputs > 0, "'num_outputs' cannot be less than or equal to zero, please enter a valid integer."
        assert min_hidden_layers < max_hidden_layers, "'min_hidden_layers' must be less than 'max_hidden_layers."
        assert min_channel_size < max_channel_size, "'min_channel_size' must be less than 'max_channel_size'."
        if block_type != "Conv2d":
            raise NotImplementedError("only Conv2d encoders are implemented in the CUDA kernels")
        if layer_norm:
            raise NotImplementedError("BatchNorm inside the CNN encoder is not implemented in the CUDA kernels")
        assert len(input_shape) == 3, f"For Conv2d, input_shape should be (channels, height, width), got {input_shape}"
        self.input_shape = tuple(int(d) for d in input_shape)
        self.num_outputs = num_outputs
        self.channel_size = list(channel_size)
        self.kernel_size = [int(k[-1]) if isinstance(k, (tuple, list)) else int(k) for k in kernel_size]
        self.stride_size = [int(s[-1]) if isinstance(s, (tuple, list)) else int(s) for s in stride_size]
        self.sample_input = None
        self.block_type = block_type
        self.activation, self.output_activation = activation, output_activation
        self.min_hidden_layers, self.max_hidden_layers = min_hidden_layers, max_hidden_layers
        self.min_channel_size, self.max_channel_size = min_channel_size, max_channel_size
        self.layer_norm, self.init_layers = layer_norm, init_layers
        self.name = name
        self._check_geometry()

    # -- geometry ------------------------------------------------------------------------------
    def _spatial(self) -> list[tuple[int, int]]:
        h, w = self.input_shape[-2:]
        out = []
        for k, s in zip(self.kernel_size, self.stride_size):
            h, w = _conv_out(h, k, s), _conv_out(w, k, s)
            out.append((h, w))
        return out

    def _check_geometry(self) -> None:
        for h, w in self._spatial():
            if h < 1 or w < 1:
                raise ValueError("convolution stack collapses the image to zero size")

    @property
    def cnn_output_size(self) -> tuple:
        h, w = self._spatial()[-1]
        return (1, self.channel_size[-1], h, w)

    def calc_max_kernel_sizes(self) -> list[int]:
        """cnn.py:110-148: clamp(0.25 * min(out_h, out_w), 1, 9) per layer."""
        out = []
        for h, w in self._spatial():
            m = int(min(h, w) * 0.25)
            out.append(1 if m <= 0 else min(m, 9))
        return out

    # -- mutations -----------------------------------------------------------------------------
    @mutation(MutationType.LAYER)
    def add_layer(self):
        """cnn.py:582-656."""
        h, w = self._spatial()[-1]
        max_k = self.calc_max_kernel_sizes()
        if (len(self.channel_size) < self.max_hidden_layers and h > 2 and w > 2 and max_k and max_k[-1] > 2):
            l_in = w
            if l_in < 2:
                return self.add_channel()
            k_new = int(self.rng.integers(2, l_in + 1))
            k_new = min(k_new, h, w)
            max_s = min(h, w) - k_new + 1
            if max_s < 1:
                return self.add_channel()
            s_new = int(self.rng.integers(1, max_s + 1))
            self.channel_size += [self.channel_size[-1]]
            self.kernel_size += [k_new]
            self.stride_size += [s_new]
            return None
        return self.add_channel()

    @mutation(MutationType.LAYER, shrink_params=True)
    def remove_layer(self):
        """cnn.py:658-672."""
        if len(self.channel_size) > self.min_hidden_layers:
            self.channel_size = self.channel_size[:-1]
            self.kernel_size = self.kernel_size[:-1]
            self.stride_size = self.stride_size[:-1]
            return None
        return self.add_channel()

    @mutation(MutationType.NODE)
    def change_kernel(self, kernel_size: int | None = None, hidden_layer: int | None = None):
        """cnn.py:674-704: re-draw the kernel of a layer in [1, min(4, n_layers))."""
        if len(self.channel_size) > 1:
            if hidden_layer is None:
                hidden_layer = int(self.rng.integers(1, min(4, len(self.channel_size))))
            if kernel_size is None:
                max_k = self.calc_max_kernel_sizes()[hidden_layer]
                kernel_size = int(self.rng.integers(1, max_k + 1))
            old = self.kernel_size[hidden_layer]
            self.kernel_size[hidden_layer] = int(kernel_size)
            try:
                self._check_geometry()
            except ValueError:
                self.kernel_size[hidden_layer] = old
            return {"hidden_layer": hidden_layer, "kernel_size": self.kernel_size[hidden_layer]}
        return self.add_layer()

    @mutation(MutationType.NODE)
    def add_channel(self, hidden_layer: int | None = None, numb_new_channels: int | None = None) -> dict:
        """cnn.py:706-734."""
        if hidden_layer is None:
            hidden_layer = int(self.rng.integers(0, len(self.channel_size)))
        else:
            hidden_layer = min(hidden_layer, len(self.channel_size) - 1)
        if numb_new_channels is None:
            numb_new_channels = int(self.rng.choice([8, 16, 32]))
        if self.channel_size[hidden_layer] + numb_new_channels <= self.max_channel_size:
            self.channel_size[hidden_layer] += numb_new_channels
        return {"hidden_layer": hidden_layer, "numb_new_channels": numb_new_channels}

    @mutation(MutationType.NODE, shrink_params=True)
    def remove_channel(self, hidden_layer: int | None = None, numb_new_channels: int | None = None) -> dict:
        """cnn.py:736-765."""
        if hidden_layer is None:
            hidden_layer = int(self.rng.integers(0, len(self.channel_size)))
        else:
            hidden_layer = min(hidden_layer, len(self.channel_size) - 1)
        if numb_new_channels is None:
            numb_new_channels = int(self.rng.choice([8, 16, 32]))
        if self.channel_size[hidden_layer] - numb_new_channels >= self.min_channel_size:
            self.channel_size[hidden_layer] -= numb_new_channels
        else:
            numb_new_channels = 0
        return {"hidden_layer": hidden_layer, "numb_new_channels": numb_new_channels}
