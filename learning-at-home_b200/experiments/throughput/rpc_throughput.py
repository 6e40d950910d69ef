"""
``torch.distributed.rpc`` baseline of the throughput experiment (CLI parity with /root/reference/experiments/throughput/
rpc_throughput.py: rank 0 drives, ranks 1..W-1 host the layers round-robin, CPU tensors on the wire, synchronous calls).

    MASTER_ADDR=127.0.0.1 MASTER_PORT=29500 python -m lah_b200.experiments.throughput.rpc_throughput \
        --rank R --world-size W --block-type ffn

Every worker keeps a registry {layer index -> block}; the driver's chain is a list of hops whose callable is one blocking
RPC to the owner of that layer (harness.Chain does the rest).
"""
import os
from argparse import ArgumentParser

import torch

from ...models.layers import name_to_block, name_to_input
from .harness import Chain, Hop, add_common_flags, meter_from_args

_LAYERS = {}   # per-process registry of the layers this worker hosts


def _host_layer(index: int, block_type: str, hid_dim: int) -> int:
    device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    _LAYERS[index] = (name_to_block[block_type](hid_dim).to(device).eval(), device)
    return index


@torch.no_grad()
def _apply_layer(index: int, x: torch.Tensor) -> torch.Tensor:
    block, device = _LAYERS[index]
    return block(x.to(device)).cpu()


def worker_name(rank: int) -> str:
    return f"worker{rank}"


def build_remote_chain(args) -> Chain:
    import torch.distributed.rpc as rpc
    hosts = [worker_name(r) for r in range(1, args.world_size)]
    total = args.layers_per_gpu * len(hosts)
    owners = [hosts[i % len(hosts)] for i in range(total)]
    for f in [rpc.rpc_async(owner, _host_layer, args=(i, args.block_type, args.hid_dim)) for i, owner in enumerate(owners)]:
        f.wait()
    return Chain([Hop(lambda x, i=i, owner=owner: rpc.rpc_sync(owner, _apply_layer, args=(i, x))) for i, owner in enumerate(owners)])


def run(args, printer=print):
    import torch.distributed.rpc as rpc
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    rpc.init_rpc(worker_name(args.rank), rank=args.rank, world_size=args.world_size)
    result = None
    try:
        if args.rank == 0:
            torch.manual_seed(0)
            chain = build_remote_chain(args)
            x = name_to_input[args.block_type](args.batch_size, args.hid_dim).normal_()
            m = meter_from_args(args).measure(lambda: chain(x))
            result = (m.throughput, m.throughput_std, m.latency, m.latency_std)
            printer(f"ModelParallel:\t{m.throughput:.2f}±{m.throughput_std:.2f}")
    finally:
        rpc.shutdown()
    return result


def make_parser():
    parser = add_common_flags(ArgumentParser(description=__doc__), pings=False)
    parser.add_argument("--rank", type=int, required=True)
    parser.add_argument("--world-size", type=int, required=True)
    return parser


if __name__ == "__main__":
    run(make_parser().parse_args())
