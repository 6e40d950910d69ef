"""
In-process baselines of the throughput experiment (CLI parity:
/root/reference/experiments/throughput/baseline_throughput.py:13-101):

* ModelParallelNetwork      — layers_per_gpu consecutive blocks per device, one hop between devices, no micro-batching
* DummyCrowdsourcedNetwork  — blocks round-robined over devices with sleep(ping * Weibull(1)) after every layer

    python -m lah_b200.experiments.throughput.baseline_throughput --block-type ffn --gpus 0 1 --layers-per-gpu 8
(`--gpus -1` runs on the CPU.)
"""
import time
from argparse import ArgumentParser
from functools import partial
from itertools import chain, repeat

import numpy as np
import torch
import torch.nn as nn

from ...models.layers import name_to_block, name_to_input


def _device(g):
    return torch.device("cuda", g) if g >= 0 and torch.cuda.is_available() else torch.device("cpu")


class ModelParallelNetwork(nn.Module):
    def __init__(self, hid_dim, block_factory, gpus, layers_per_gpu):
        super().__init__()
        self.devices = [_device(g) for g in gpus]
        self.blocks = nn.ModuleList([nn.Sequential(*(block_factory(hid_dim) for _ in range(layers_per_gpu))).to(dev)
                                     for dev in self.devices])

    def forward(self, x, ping=None):
        for dev, stage in zip(self.devices, self.blocks):
            x = stage(x.to(dev, non_blocking=True))
        return x


class DummyCrowdsourcedNetwork(nn.Module):
    def __init__(self, hid_dim, block_factory, gpus, layers_per_gpu):
        super().__init__()
        self.devices = [_device(g) for g in chain.from_iterable(repeat(gpus, layers_per_gpu))]
        self.layers = nn.ModuleList([block_factory(hid_dim).to(dev) for dev in self.devices])

    def forward(self, x, ping):
        for dev, layer in zip(self.devices, self.layers):
            x = layer(x.to(dev, non_blocking=True))
            if ping:
                time.sleep(ping * np.random.weibull(1))  # emulated network lag
        return x


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def measure_perf(model_class, batches_for_latency, batches_for_throughput, throughput_runs, ping, input_factory,
                 batch_size, hid_dim, **kwargs):
    model = model_class(hid_dim=hid_dim, **kwargs).eval()
    pin = torch.cuda.is_available()
    z = input_factory(batch_size, hid_dim).normal_()
    out_buf = input_factory(batch_size, hid_dim)
    if pin:
        z, out_buf = z.pin_memory(), out_buf.pin_memory()
    time_per_batch, throughputs = [], []
    with torch.no_grad():
        for _ in range(batches_for_latency + 1):  # latency: time to obtain the result of one batch
            start = time.time()
            out_buf.copy_(model(z, ping=ping), non_blocking=True)
            _sync()
            time_per_batch.append(time.time() - start)
        for _ in range(throughput_runs):  # throughput: samples/s with asynchronous results
            start = time.time()
            for _ in range(batches_for_throughput):
                out_buf.copy_(model(z, ping=ping), non_blocking=True)
            _sync()
            throughputs.append(batch_size * batches_for_throughput / (time.time() - start))
    lat = time_per_batch[1:]
    std = lambda v: float(np.std(v, ddof=1)) if len(v) > 1 else 0.0  # noqa: E731
    return float(np.mean(lat)), std(lat), float(np.mean(throughputs)), std(throughputs)


def run(args, printer=print):
    np.random.seed(0)
    torch.manual_seed(0)
    measure = partial(measure_perf, batches_for_latency=args.batches_for_latency,
                      batches_for_throughput=args.batches_for_throughput, throughput_runs=args.throughput_runs,
                      gpus=args.gpus, layers_per_gpu=args.layers_per_gpu, hid_dim=args.hid_dim,
                      block_factory=name_to_block[args.block_type], batch_size=args.batch_size,
                      input_factory=name_to_input[args.block_type])
    rows = []
    lat, lat_std, thr, thr_std = measure(ModelParallelNetwork, ping=0)
    rows.append(("fast", 0.0, lat, thr))
    printer(f"ModelParallel (fast, ping=0.00):\t{lat:.2f}±{lat_std:.2f}\t{thr:.2f}±{thr_std:.2f}")
    for ping in np.linspace(0, args.max_ping, args.linspace_points):
        lat, lat_std, thr, thr_std = measure(DummyCrowdsourcedNetwork, ping=ping)
        rows.append(("slow", float(ping), lat, thr))
        printer(f"ModelParallel (slow, ping={ping:.2f}):\t{lat:.2f}±{lat_std:.2f}\t{thr:.2f}±{thr_std:.2f}")
    return rows


def make_parser():
    parser = ArgumentParser()
    parser.add_argument("--hid-dim", type=int, default=1024)
    parser.add_argument("--batches-for-latency", type=int, default=10)
    parser.add_argument("--batches-for-throughput", type=int, default=100)
    parser.add_argument("--batch-size", type=int, default=2048)
    parser.add_argument("--throughput-runs", type=int, default=10)
    parser.add_argument("--max-ping", type=float, default=0.2)
    parser.add_argument("--linspace-points", type=int, default=10)
    parser.add_argument("--gpus", type=int, nargs="+", required=True)
    parser.add_argument("--layers-per-gpu", type=int, default=56)
    parser.add_argument("--block-type", choices=name_to_block.keys(), required=True)
    return parser


if __name__ == "__main__":
    run(make_parser().parse_args())
