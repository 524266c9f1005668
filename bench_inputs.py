"""Seeded synthetic inputs of bench.py's native arm (SURVEY.md 8(d) "Synthetic inputs"): the images and the random-init
regressor weights.  Not part of the oracle and not part of the product: bench.py's native arm takes its inputs from here so
that it touches oracle/ only in the cpu_baseline leg and the reference arm; tests/test_bench_contract_cpu.py checks that these
generators and the oracle's produce the same bits (so the parity tests and the bench talk about the same inputs)."""
from __future__ import annotations

from collections import OrderedDict

import torch


def synthetic_image(index: int, h: int, w: int) -> torch.Tensor:
    """Image `index`: 0.05 + 0.9 * U[0, 1) from its own generator seeded 1000 + index (independent of sharding), [3, h, w]."""
    g = torch.Generator().manual_seed(1000 + index)
    return 0.05 + 0.9 * torch.rand(3, h, w, generator=g)


def make_regressor_state_dict(num_classes: int = 4) -> "OrderedDict[str, torch.Tensor]":
    """torchvision resnet50 under torch.manual_seed(0) with an fc of `num_classes` outputs; then, from a generator seeded 1,
    per BatchNorm2d in named_modules() order: running_mean = 0.1 * randn, running_var = 0.5 + rand, bias = 0.1 * randn; every
    bn3 weight = 0.4 (plain init saturates the sigmoid and zero_init_residual would kill the bottleneck branches)."""
    from torchvision import models
    torch.manual_seed(0)
    net = models.resnet50()
    net.fc = torch.nn.Linear(net.fc.in_features, num_classes)
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, mod in net.named_modules():
            if not isinstance(mod, torch.nn.BatchNorm2d):
                continue
            mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=gen))
            mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=gen))
            mod.bias.copy_(0.1 * torch.randn(mod.num_features, generator=gen))
            if name.endswith("bn3"):
                mod.weight.fill_(0.4)
    return OrderedDict((k, v.detach().clone()) for k, v in net.state_dict().items())
