"""
Multi-trainer client of the throughput experiment (CLI parity:
/root/reference/experiments/throughput/throughput_client.py:18-85; same flags and the same metric definitions:
latency = mean wall time per batch excluding the first; throughput = jobs * batch * (batches + 1) / wall).

Experts of different hosts are interleaved into one chain; after every layer the trainer sleeps ping * Weibull(1) to
emulate network latency; `--jobs` concurrent trainers (threads here; the RPCs release the GIL) play the role of
pipelining.  Forward only, like the reference.

    python -m lah_b200.experiments.throughput.throughput_client -j 64 --hosts 127.0.0.1:8080 127.0.0.1:8081 \
        --block-type ffn --layers-per-gpu 56
"""
from argparse import ArgumentParser
from concurrent.futures import ThreadPoolExecutor
from functools import partial
from itertools import chain
from time import sleep, time

import numpy as np
import torch
import torch.nn as nn

from ... import RemoteExpert
from ...models.layers import name_to_block, name_to_input


class ExpertsWithLatency(nn.Module):
    def __init__(self, experts):
        super().__init__()
        self.experts = nn.Sequential(*experts)

    def forward(self, x, ping):
        for layer in self.experts:
            x = layer(x)
            if ping:
                sleep(ping * np.random.weibull(1))
        return x


@torch.no_grad()
def measure_perf(ping, model, x, num_batches):
    latencies = []
    for _ in range(num_batches + 1):
        start = time()
        model(x, ping=ping)
        latencies.append(time() - start)
    return latencies[1:]


def build_chain(hosts, layers_per_gpu):
    per_host = []
    for address in hosts:
        host, port = address.split(":")
        per_host.append([RemoteExpert(f"expert{i}", host=host, port=int(port)) for i in range(layers_per_gpu)])
    return ExpertsWithLatency(list(chain.from_iterable(zip(*per_host))))  # interleave across hosts


def run(args, printer=print):
    np.random.seed(0)
    torch.manual_seed(0)
    model = build_chain(args.hosts, args.layers_per_gpu)
    x = name_to_input[args.block_type](args.batch_size, args.hid_dim).normal_()
    measure = partial(measure_perf, model=model, x=x, num_batches=args.batches_for_throughput)
    results = []
    with ThreadPoolExecutor(args.jobs) as pool:
        for ping in np.linspace(0, args.max_ping, args.linspace_points):
            latencies = measure_perf(ping, model, x, args.batches_for_latency)
            throughputs = []
            for _ in range(args.throughput_runs):
                start = time()
                list(pool.map(measure, [ping] * args.jobs))
                throughputs.append(args.jobs * args.batch_size * (args.batches_for_throughput + 1) / (time() - start))
            row = dict(ping=float(ping), latency=float(np.mean(latencies)),
                       latency_std=float(np.std(latencies, ddof=1)) if len(latencies) > 1 else 0.0,
                       throughput=float(np.mean(throughputs)),
                       throughput_std=float(np.std(throughputs, ddof=1)) if len(throughputs) > 1 else 0.0)
            results.append(row)
            printer(f"ModelParallel (ours, ping={ping:.2f}):\t{row['latency']:.2f}±{row['latency_std']:.2f}\t"
                    f"{row['throughput']:.2f}±{row['throughput_std']:.2f}")
    return results


def make_parser():
    parser = ArgumentParser()
    parser.add_argument("-j", "--jobs", type=int, required=True)
    parser.add_argument("--hosts", nargs="+", required=True)
    parser.add_argument("--hid-dim", type=int, default=1024)
    parser.add_argument("--batches-for-latency", type=int, default=10)
    parser.add_argument("--batches-for-throughput", type=int, default=100)
    parser.add_argument("--throughput-runs", type=int, default=10)
    parser.add_argument("--batch-size", type=int, default=2048)
    parser.add_argument("--linspace-points", type=int, default=10)
    parser.add_argument("--layers-per-gpu", type=int, default=56)
    parser.add_argument("--block-type", choices=name_to_block.keys(), required=True)
    parser.add_argument("--max-ping", type=float, default=0.2)
    return parser


if __name__ == "__main__":
    run(make_parser().parse_args())
