"""
Harness (NOT product code): the reference's OWN RPC stack on this box, for the throughput experiment.

Runs the UNMODIFIED reference package from baseline/_ref through its public API exactly like
experiments/throughput/throughput_server.py + throughput_client.py do: one `lib.TesseractServer(None, experts, ...)` per
device hosting `layers_per_gpu` jit-scripted experts behind `lib.ExpertBackend`s, and `jobs` client processes that push
batches through the chain of `lib.RemoteExpert`s (forward only, ping = 0).  Throughput = jobs * batch * (batches + 1) / wall
(throughput_client.py:64).  Needs the two import shims in baseline/shims (prefetch_generator, kademlia) and
TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1.

    python baseline/ref_throughput.py --block-type ffn --layers-per-gpu 4 --jobs 8 --device cuda:0
"""
import argparse
import json
import multiprocessing as mp
import os
import signal
import socket
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
sys.path.insert(0, os.path.join(ROOT, "baseline", "shims"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))


def server_main(args, port, device):
    os.setsid()  # own process group: the parent kills the whole tree (handlers, pools) afterwards
    import multiprocessing.managers
    import torch
    import lib
    from experiments.throughput.layers import name_to_block
    inp_shape = (args.hid_dim,) if args.block_type == "ffn" else (512, args.hid_dim)
    with multiprocessing.managers.SharedMemoryManager() as shm_manager, multiprocessing.Manager() as hdr_manager:
        array_headers = hdr_manager.dict()
        experts = {}
        for i in range(args.layers_per_gpu):
            expert = torch.jit.script(name_to_block[args.block_type](args.hid_dim))
            experts[f"expert{i}"] = lib.ExpertBackend(name=f"expert{i}", expert=expert,
                                                      opt=torch.optim.Adam(expert.parameters()),
                                                      args_schema=(lib.BatchTensorProto(*inp_shape),),
                                                      outputs_schema=lib.BatchTensorProto(*inp_shape),
                                                      max_batch_size=args.max_batch_size, shm_manager=shm_manager,
                                                      array_headers=array_headers, pool_size=8)
        lib.TesseractServer(None, experts, port=port, conn_handler_processes=args.handler_processes, sender_threads=4,
                            device=torch.device(device), start=True)


def client_job(job):
    args, ports, num_batches = job
    import torch
    import lib.client
    from itertools import chain
    from experiments.throughput.layers import name_to_input
    per_host = [[lib.RemoteExpert(f"expert{i}", host="127.0.0.1", port=p) for i in range(args.layers_per_gpu)] for p in ports]
    experts = list(chain.from_iterable(zip(*per_host)))
    x = name_to_input[args.block_type](args.batch_size, args.hid_dim).normal_()
    with torch.no_grad():
        for _ in range(num_batches + 1):
            y = x
            for layer in experts:
                y = layer(y)
    return float(y.float().abs().mean())


def wait_port(port, timeout):
    t0 = time.time()
    while time.time() - t0 < timeout:
        with socket.socket() as s:
            if s.connect_ex(("127.0.0.1", port)) == 0:
                return True
        time.sleep(0.5)
    return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--block-type", choices=["ffn", "transformer"], default="ffn")
    ap.add_argument("--hid-dim", type=int, default=1024)
    ap.add_argument("--layers-per-gpu", type=int, default=56)
    ap.add_argument("--devices", nargs="+", default=["cuda:0"])
    ap.add_argument("-j", "--jobs", type=int, default=64)
    ap.add_argument("-a", "--handler-processes", type=int, default=16)
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--max-batch-size", type=int, default=2048)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--base-port", type=int, default=18080)
    ap.add_argument("--startup-timeout", type=float, default=180.0)
    args = ap.parse_args()
    args.batch_size = args.batch_size or (2048 if args.block_type == "ffn" else 4)
    mp.set_start_method("fork")
    ports = [args.base_port + i for i in range(len(args.devices))]
    servers = [mp.Process(target=server_main, args=(args, port, dev), daemon=False) for port, dev in zip(ports, args.devices)]
    out = dict(impl="reference", what="reference RPC stack (TesseractServer + RemoteExpert over TCP), forward chain",
               block_type=args.block_type, layers_total=args.layers_per_gpu * len(ports), jobs=args.jobs,
               batch_size=args.batch_size, handler_processes=args.handler_processes, devices=args.devices)
    try:
        for s in servers:
            s.start()
        if not all(wait_port(p, args.startup_timeout) for p in ports):
            out["unavailable"] = "reference server did not come up"
        else:
            time.sleep(2.0)
            with mp.Pool(args.jobs) as pool:
                pool.map(client_job, [(args, ports, 0)] * args.jobs)      # warm-up: info RPCs, first batches
                t0 = time.time()
                pool.map(client_job, [(args, ports, args.batches)] * args.jobs)
                wall = time.time() - t0
            out.update(value=args.jobs * args.batch_size * (args.batches + 1) / wall, unit="samples/s", wall_s=wall)
    except Exception as e:  # noqa
        out["unavailable"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        for s in servers:
            if s.pid:
                try:
                    os.killpg(s.pid, signal.SIGKILL)   # exactly the process groups we created
                except ProcessLookupError:
                    pass
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
