"""
Harness (NOT product code) that drives the UNMODIFIED reference emulator from baseline/_ref for `bench.py --impl reference`.

Everything model-related is the reference's own code path:
  * `experiments.convergence.dmoe_emulator.EmulatedDMoE` / `get_non_expert_params`     (stock)
  * `experiments.convergence.faulty_dmoe_emulator.EmulatedFaultyDMoE`                  (stock, failure_rate > 0)
  * `experiments.throughput.layers.FeedforwardBlock`                                   (stock expert)
The training loop below is the notebook's trainer loop (cell 3 of convergence_mnist_..._dmoe64x4.ipynb) for ONE
synchronous trainer without the artificial latency sleep: forward, cross-entropy, backward, Adam(lr=1e-3, amsgrad=True)
on the non-expert parameters; experts step themselves inside forward (update_every_inputs=batch_size,
update_every_steps=10, notebook cell 2).
"""
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F


def build_reference_trainer(hidden=512, num_experts=64, num_active=4, num_layers=4, batch_size=64, failure_rate=0.0,
                            seed=1337, in_features=784, num_classes=10, update_every_steps=10):
    from experiments.convergence.dmoe_emulator import EmulatedDMoE, get_non_expert_params
    from experiments.convergence.faulty_dmoe_emulator import EmulatedFaultyDMoE
    from experiments.convergence import faulty_dmoe_emulator
    from experiments.throughput.layers import FeedforwardBlock

    torch.manual_seed(seed)
    device = torch.device("cuda", torch.cuda.current_device())
    Optimizer = partial(torch.optim.Adam, lr=1e-3, amsgrad=True)
    if failure_rate > 0:
        make = lambda: EmulatedFaultyDMoE(hidden, num_experts=num_experts, num_active=num_active,  # noqa
                                          update_every_inputs=batch_size, update_every_steps=update_every_steps,
                                          failure_rate=failure_rate, Expert=FeedforwardBlock, Optimizer=Optimizer)
        non_expert = faulty_dmoe_emulator.get_non_expert_params
    else:
        make = lambda: EmulatedDMoE(hidden, num_experts=num_experts, num_active=num_active,  # noqa
                                    update_every_inputs=batch_size, update_every_steps=update_every_steps,
                                    Expert=FeedforwardBlock, Optimizer=Optimizer)
        non_expert = get_non_expert_params
    model = nn.Sequential(nn.Linear(in_features, hidden), *(make() for _ in range(num_layers)),
                          nn.LayerNorm(hidden), nn.Linear(hidden, num_classes)).to(device)
    opt = Optimizer(non_expert(model))
    # same synthetic MNIST-shaped learnable task as the other arm (bench.py synthetic_mnist; duplicated here so that this
    # harness imports nothing of the product)
    gen = torch.Generator().manual_seed(4242)
    protos = torch.randn(num_classes, in_features, generator=gen)
    gen = torch.Generator().manual_seed(seed)
    xs_host, ys_host = [], []
    for _ in range(4):
        yb = torch.randint(0, num_classes, (batch_size,), generator=gen)
        xs_host.append((protos[yb] + 3.0 * torch.randn(batch_size, in_features, generator=gen)).pin_memory())
        ys_host.append(yb.pin_memory())
    xs_dev = [x.to(device) for x in xs_host]
    ys_dev = [y.to(device) for y in ys_host]
    loss_host = torch.empty(1).pin_memory()

    def step(i, e2e=False):
        if e2e:
            xb = xs_host[i % 4].to(device, non_blocking=True)
            yb = ys_host[i % 4].to(device, non_blocking=True)
        else:
            xb, yb = xs_dev[i % 4], ys_dev[i % 4]
        model.train(True)
        logits = model(xb)
        loss = F.cross_entropy(logits, yb)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if e2e:
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(loss_host[0])
        return loss

    info = (f"reference EmulatedDMoE: Linear(784,{hidden}) -> {num_layers} x EmulatedDMoE[{num_experts} experts "
            f"FeedforwardBlock({hidden}), top-{num_active}] -> LayerNorm -> Linear({hidden},10); fp32; "
            f"stock per-sample expert loop; batch {batch_size}")
    return step, info
