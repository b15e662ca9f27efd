"""Time the priority write-back alone (CUDA events, back-to-back launches): python tools/bench_tree.py"""
import torch
from agilerl_b200.components import PrioritizedReplayBuffer, MultiStepReplayBuffer

dev = torch.device("cuda:0")
mem = PrioritizedReplayBuffer(100_000, 0.6, device=dev)
cap = mem._cap
g = torch.Generator(device=dev).manual_seed(0)
mem._size = 100_000
for s in range(0, 100_000, 4096):
    n = min(4096, 100_000 - s)
    mem.update_priorities_device(torch.arange(s, s + n, device=dev), torch.rand(n, device=dev, generator=g) + 1e-3)
for B in (32, 256, 512):
    idx = torch.randint(0, 100_000, (B,), device=dev, generator=g)
    pri = torch.rand(B, device=dev, generator=g) + 1e-3
    for _ in range(10):
        mem.update_priorities_device(idx, pri)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        mem.update_priorities_device(idx, pri)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per tree write-back")
