"""
Multi-trainer client of the throughput experiment (CLI parity with /root/reference/experiments/throughput/
throughput_client.py: same flags, same metrics — latency = mean wall time per batch without the first, throughput =
jobs * batch * (batches + 1) / wall — and the same printed lines).

The chain visits the experts of all hosts interleaved (expert i of host 0, expert i of host 1, ..., expert i+1 of host 0,
...), every hop is one blocking ``RemoteExpert`` call followed by the emulated lag sleep(ping * Weibull(1)); `--jobs`
concurrent trainers keep the servers busy (threads: the socket I/O releases the GIL).  Forward only, like the reference.

    python -m lah_b200.experiments.throughput.throughput_client -j 64 --hosts 127.0.0.1:8080 127.0.0.1:8081 \
        --block-type ffn --layers-per-gpu 56
"""
from argparse import ArgumentParser

import numpy as np
import torch

from ... import RemoteExpert
from ...models.layers import name_to_input
from .harness import Chain, Hop, LatencyModel, add_common_flags, meter_from_args, ping_grid


def parse_host(address: str):
    host, _, port = address.rpartition(":")
    return host, int(port)


def build_chain(hosts, layers_per_gpu) -> Chain:
    endpoints = [parse_host(a) for a in hosts]
    hops = []
    for layer in range(layers_per_gpu):          # interleave: consecutive hops always change host
        for host, port in endpoints:
            hops.append(Hop(RemoteExpert(f"expert{layer}", host=host, port=port)))
    return Chain(hops)


def run(args, printer=print):
    np.random.seed(0)
    torch.manual_seed(0)
    chain = build_chain(args.hosts, args.layers_per_gpu)
    x = name_to_input[args.block_type](args.batch_size, args.hid_dim).normal_()
    meter = meter_from_args(args)
    results = []
    for ping in ping_grid(args):
        lag = LatencyModel(ping) if ping else None
        m = meter.measure(lambda: chain(x, latency=lag), concurrency=args.jobs, count_warm_batch=True)
        row = dict(ping=ping, latency=m.latency, latency_std=m.latency_std, throughput=m.throughput,
                   throughput_std=m.throughput_std)
        results.append(row)
        printer(m.line(f"ModelParallel (ours, ping={ping:.2f})"))
    return results


def make_parser():
    parser = add_common_flags(ArgumentParser(description=__doc__))
    parser.add_argument("-j", "--jobs", type=int, required=True)
    parser.add_argument("--hosts", nargs="+", required=True)
    return parser


if __name__ == "__main__":
    run(make_parser().parse_args())
