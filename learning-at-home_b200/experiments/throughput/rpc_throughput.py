"""
torch.distributed.rpc baseline (CLI parity: /root/reference/experiments/throughput/rpc_throughput.py:16-95): one
BlockWorker per layer round-robined over world_size-1 workers, a synchronous rpc chain with CPU tensors on the wire.

    MASTER_ADDR=127.0.0.1 MASTER_PORT=29500 python -m lah_b200.experiments.throughput.rpc_throughput \
        --rank R --world-size W --block-type ffn      (rank 0 drives, ranks 1..W-1 host layers)
"""
import time
from argparse import ArgumentParser
from itertools import chain, repeat

import numpy as np
import torch
import torch.nn as nn

from ...models.layers import name_to_block, name_to_input


class BlockWorker(nn.Module):
    def __init__(self, hid_dim, block_type):
        super().__init__()
        self.device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.block = name_to_block[block_type](hid_dim).to(self.device).eval()

    @torch.no_grad()
    def forward(self, x):
        return self.block(x.to(self.device)).cpu()


def _call_method(method, rref, *args, **kwargs):
    return method(rref.local_value(), *args, **kwargs)


def _remote_method(method, rref, *args, **kwargs):
    from torch.distributed.rpc import rpc_sync
    return rpc_sync(rref.owner(), _call_method, args=[method, rref] + list(args), kwargs=kwargs)


class ModelParallelRPC(nn.Module):
    def __init__(self, hid_dim, block_type, workers, layers_per_gpu):
        super().__init__()
        import torch.distributed.rpc as rpc
        self.workers = list(chain.from_iterable(repeat(workers, layers_per_gpu)))
        self.layer_rrefs = [rpc.remote(worker, BlockWorker, args=(hid_dim, block_type)) for worker in self.workers]

    def forward(self, x):
        for rref in self.layer_rrefs:
            x = _remote_method(BlockWorker.forward, rref, x)
        return x


def measure_perf(model, batches_for_throughput, throughput_runs, input_factory, batch_size, hid_dim):
    z = input_factory(batch_size, hid_dim).normal_()
    throughputs = []
    with torch.no_grad():
        for _ in range(throughput_runs):
            start = time.time()
            for _ in range(batches_for_throughput):
                model(z)
            throughputs.append(batch_size * batches_for_throughput / (time.time() - start))
    return float(np.mean(throughputs)), float(np.std(throughputs, ddof=1)) if len(throughputs) > 1 else 0.0


def run(args, printer=print):
    import torch.distributed.rpc as rpc
    rpc.init_rpc(f"worker{args.rank}", rank=args.rank, world_size=args.world_size)
    result = None
    if args.rank == 0:
        np.random.seed(0)
        torch.manual_seed(0)
        model = ModelParallelRPC(args.hid_dim, args.block_type, [f"worker{r}" for r in range(1, args.world_size)],
                                 args.layers_per_gpu)
        result = measure_perf(model, args.batches_for_throughput, args.throughput_runs, name_to_input[args.block_type],
                              args.batch_size, args.hid_dim)
        printer(f"ModelParallel:\t{result[0]:.2f}±{result[1]:.2f}")
    rpc.shutdown()
    return result


def make_parser():
    parser = ArgumentParser()
    parser.add_argument("--hid-dim", type=int, default=1024)
    parser.add_argument("--batches-for-latency", type=int, default=10)
    parser.add_argument("--batches-for-throughput", type=int, default=100)
    parser.add_argument("--batch-size", type=int, default=2048)
    parser.add_argument("--throughput-runs", type=int, default=10)
    parser.add_argument("--rank", type=int, required=True)
    parser.add_argument("--world-size", type=int, required=True)
    parser.add_argument("--layers-per-gpu", type=int, default=56)
    parser.add_argument("--block-type", choices=name_to_block.keys(), required=True)
    return parser


if __name__ == "__main__":
    run(make_parser().parse_args())
