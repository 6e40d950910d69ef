"""
Measurement harness shared by the throughput CLIs (throughput_client / baseline_throughput / rpc_throughput).

The reference repeats a ``measure_perf`` loop in each script (experiments/throughput/*.py); here the three scripts only
describe WHERE each layer runs and HOW a layer is called, and this module owns the common parts:

* ``Hop``              one step of a chain: a callable plus the device its input must live on
* ``LatencyModel``     the emulated network lag of the experiment: sleep(ping * Weibull(1)) after a hop (README.md:40-42)
* ``Chain``            runs a list of hops under ``torch.no_grad`` with an optional latency model
* ``Meter``            latency (mean wall time per batch, first batch excluded) and throughput (samples / wall second with
                       ``concurrency`` independent drivers) — the metric definitions of SURVEY.md §6
* ``add_common_flags`` the flags every script of the experiment accepts
"""
import statistics
import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch


@dataclass
class Hop:
    fn: Callable
    device: Optional[torch.device] = None   # None: the callable takes the tensor wherever it is (remote experts)


class LatencyModel:
    def __init__(self, ping: float, seed: Optional[int] = None):
        self.ping = float(ping)
        self._rng = np.random.default_rng(seed)

    def wait(self):
        if self.ping > 0:
            time.sleep(self.ping * float(self._rng.weibull(1)))


class Chain:
    def __init__(self, hops: Sequence[Hop]):
        self.hops = list(hops)

    def modules(self):
        return [h.fn for h in self.hops if isinstance(h.fn, torch.nn.Module)]

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, latency: Optional[LatencyModel] = None, lag_every_hop: bool = True):
        for hop in self.hops:
            if hop.device is not None and x.device != hop.device:
                x = x.to(hop.device, non_blocking=True)
            x = hop.fn(x)
            if latency is not None and lag_every_hop:
                latency.wait()
        return x


@dataclass
class Measurement:
    latency: float
    latency_std: float
    throughput: float
    throughput_std: float

    def line(self, label: str) -> str:
        return f"{label}:\t{self.latency:.2f}±{self.latency_std:.2f}\t{self.throughput:.2f}±{self.throughput_std:.2f}"


def _spread(values: List[float]) -> float:
    return statistics.stdev(values) if len(values) > 1 else 0.0


def _device_sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


class Meter:
    def __init__(self, batch_size: int, batches_for_latency: int, batches_for_throughput: int, throughput_runs: int):
        self.batch_size, self.n_lat, self.n_thr, self.runs = batch_size, batches_for_latency, batches_for_throughput, throughput_runs

    def latency(self, one_batch: Callable[[], None]) -> List[float]:
        """wall seconds of `n_lat` batches after one untimed batch (result available = device synchronised)"""
        one_batch()
        _device_sync()
        out = []
        for _ in range(self.n_lat):
            t0 = time.perf_counter()
            one_batch()
            _device_sync()
            out.append(time.perf_counter() - t0)
        return out

    def throughput(self, one_batch: Callable[[], None], concurrency: int = 1, count_warm_batch: bool = False) -> List[float]:
        """samples / second of `runs` runs; every one of `concurrency` drivers pushes `n_thr` batches (+1 when
        ``count_warm_batch``: the reference's client counts the discarded first batch of its latency loop, throughput_client.py:64)"""
        per_driver = self.n_thr + (1 if count_warm_batch else 0)

        def driver(_):
            for _ in range(per_driver):
                one_batch()

        rates = []
        pool = ThreadPoolExecutor(concurrency) if concurrency > 1 else None
        try:
            for _ in range(self.runs):
                t0 = time.perf_counter()
                if pool is None:
                    driver(0)
                else:
                    list(pool.map(driver, range(concurrency)))
                _device_sync()
                rates.append(concurrency * self.batch_size * per_driver / (time.perf_counter() - t0))
        finally:
            if pool is not None:
                pool.shutdown()
        return rates

    def measure(self, one_batch: Callable[[], None], concurrency: int = 1, count_warm_batch: bool = False) -> Measurement:
        lat = self.latency(one_batch)
        thr = self.throughput(one_batch, concurrency, count_warm_batch)
        return Measurement(statistics.fmean(lat), _spread(lat), statistics.fmean(thr), _spread(thr))


def add_common_flags(parser, *, pings: bool = True):
    from ...models.layers import name_to_block
    parser.add_argument("--hid-dim", type=int, default=1024)
    parser.add_argument("--batches-for-latency", type=int, default=10)
    parser.add_argument("--batches-for-throughput", type=int, default=100)
    parser.add_argument("--throughput-runs", type=int, default=10)
    parser.add_argument("--batch-size", type=int, default=2048)
    parser.add_argument("--layers-per-gpu", type=int, default=56)
    parser.add_argument("--block-type", choices=name_to_block.keys(), required=True)
    if pings:
        parser.add_argument("--linspace-points", type=int, default=10)
        parser.add_argument("--max-ping", type=float, default=0.2)
    return parser


def meter_from_args(args) -> Meter:
    return Meter(args.batch_size, args.batches_for_latency, args.batches_for_throughput, args.throughput_runs)


def ping_grid(args):
    return [float(p) for p in np.linspace(0, args.max_ping, args.linspace_points)]
