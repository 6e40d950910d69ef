"""
In-process baselines of the throughput experiment (CLI parity with /root/reference/experiments/throughput/
baseline_throughput.py: same flags, same two topologies, same printed lines).

  "fast"  contiguous placement: `layers_per_gpu` consecutive blocks per device, ONE device hop per stage, no emulated lag —
          naive model parallelism without micro-batches
  "slow"  round-robin placement: consecutive blocks live on different devices and every hop pays the emulated network lag
          sleep(ping * Weibull(1)) — what a crowd-sourced deployment without Learning@home's concurrency looks like

    python -m lah_b200.experiments.throughput.baseline_throughput --block-type ffn --gpus 0 1 --layers-per-gpu 8
(`--gpus -1` runs on the CPU.)  The topologies are placement functions over one generic chain (harness.Chain).
"""
from argparse import ArgumentParser

import numpy as np
import torch

from ...models.layers import name_to_block, name_to_input
from .harness import Chain, Hop, LatencyModel, add_common_flags, meter_from_args, ping_grid


def resolve_device(index: int) -> torch.device:
    return torch.device("cuda", index) if index >= 0 and torch.cuda.is_available() else torch.device("cpu")


def contiguous_placement(num_devices: int, layers_per_gpu: int):
    """layer i -> device i // layers_per_gpu"""
    return [i // layers_per_gpu for i in range(num_devices * layers_per_gpu)]


def round_robin_placement(num_devices: int, layers_per_gpu: int):
    """layer i -> device i % num_devices"""
    return [i % num_devices for i in range(num_devices * layers_per_gpu)]


def build_chain(block_type: str, hid_dim: int, gpus, placement) -> Chain:
    devices = [resolve_device(g) for g in gpus]
    hops = []
    for slot in placement:
        block = name_to_block[block_type](hid_dim).to(devices[slot]).eval()
        hops.append(Hop(block, devices[slot]))
    return Chain(hops)


def run(args, printer=print):
    np.random.seed(0)
    torch.manual_seed(0)
    meter = meter_from_args(args)
    x = name_to_input[args.block_type](args.batch_size, args.hid_dim).normal_()
    sink = torch.empty_like(x)
    if torch.cuda.is_available():
        x, sink = x.pin_memory(), sink.pin_memory()
    rows = []

    def bench(kind, placement, ping):
        chain = build_chain(args.block_type, args.hid_dim, args.gpus, placement(len(args.gpus), args.layers_per_gpu))
        lag = LatencyModel(ping, seed=0) if ping else None
        m = meter.measure(lambda: sink.copy_(chain(x, latency=lag), non_blocking=True))
        rows.append((kind, float(ping), m.latency, m.throughput))
        printer(m.line(f"ModelParallel ({kind}, ping={ping:.2f})"))

    bench("fast", contiguous_placement, 0.0)
    for ping in ping_grid(args):
        bench("slow", round_robin_placement, ping)
    return rows


def make_parser():
    parser = add_common_flags(ArgumentParser(description=__doc__))
    parser.add_argument("--gpus", type=int, nargs="+", required=True)
    return parser


if __name__ == "__main__":
    run(make_parser().parse_args())
