"""``OptimizerWrapper`` — the Adam state of one learning agent.

Mirror of agilerl/algorithms/core/optimizer_wrapper.py:63-366 for the single-network Adam case:
``zero_grad`` / ``step`` are no-ops on the host because clip + Adam + Polyak run as ONE fused
kernel inside ``learn`` (csrc/nn.cu ``adam_polyak_kernel``); what remains is the state the
reference exposes — ``lr``, ``state_dict()`` / ``load_state_dict()`` (exp_avg, exp_avg_sq, step)
— so ``clone`` and checkpoints keep working."""
from __future__ import annotations

import torch


class OptimizerWrapper:
    def __init__(self, optimizer_cls=None, networks=None, lr: float = 1e-4, network_names=None, lr_name: str = "lr",
                 optimizer_kwargs: dict | None = None, engine=None) -> None:
        self.optimizer_cls = optimizer_cls if optimizer_cls is not None else torch.optim.Adam
        if self.optimizer_cls is not torch.optim.Adam:
            raise NotImplementedError("the fused optimiser kernel implements torch.optim.Adam")
        self.networks = networks
        self.network_names = network_names or ["actor"]
        self.lr_name = lr_name
        self.lr = lr
        self.optimizer_kwargs = optimizer_kwargs or {}
        self.engine = engine
        self.optimizer = self

    def zero_grad(self) -> None:
        pass

    def step(self) -> None:
        pass

    def state_dict(self) -> dict:
        e = self.engine
        return {"step": e.step, "exp_avg": e.exp_avg.clone(), "exp_avg_sq": e.exp_avg_sq.clone(), "lr": self.lr}

    def load_state_dict(self, sd: dict, strict: bool = False) -> None:
        e = self.engine
        if sd["exp_avg"].numel() != e.exp_avg.numel():
            msg = (f"optimizer state of {sd['exp_avg'].numel()} elements does not fit a network of "
                   f"{e.exp_avg.numel()} parameters")
            if strict:
                raise ValueError(msg)
            import warnings
            warnings.warn(msg + ": Adam moments start fresh", stacklevel=2)
            return                                   # architecture changed: fresh moments (reinit_optimizers)
        if "lr" in sd and sd["lr"] is not None:
            self.lr = sd["lr"]
        e.step = int(sd["step"])
        e.exp_avg.copy_(sd["exp_avg"].to(e.exp_avg.device))
        e.exp_avg_sq.copy_(sd["exp_avg_sq"].to(e.exp_avg.device))
