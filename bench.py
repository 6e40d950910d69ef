#!/usr/bin/env python
"""
Headline benchmark: DMoE training samples/sec (whole job, device-timed, max over ranks).

Config = BASELINE.json "DMoE 256 experts (64x4) FFN block, top-4 gating, bf16": the convergence-notebook model
(Linear(784,512) -> 4 x DMoE[64 experts, top-4, FeedforwardBlock(512)] -> LayerNorm -> Linear(512,10)), full training
step = forward + backward + per-expert AMSGrad + trainer AMSGrad, synthetic MNIST-shaped data, random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch-per-gpu B]          # our engine
    python bench.py --impl reference ...                                             # unmodified reference (baseline/_ref)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference", "baseline"], default="ours",
                    help="baseline = the same step written with torch.topk + NCCL all_to_all_single + cuBLAS + torch Adam "
                         "(lah_b200/parallel/baseline.py): 'the baseline, not the product'")
    ap.add_argument("--batch-per-gpu", type=int, default=65536, help="samples per GPU per step (weak scaling)")
    ap.add_argument("--ref-batch", type=int, default=64, help="samples per step for the reference arm")
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--grid", type=int, nargs="+", default=[64])
    ap.add_argument("--gate", choices=["emulator", "product_key"], default="emulator",
                    help="emulator = EmulatedDMoE gate of the reference arm (LayerNorm + normalized keys, not trained)")
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--failure-rate", type=float, default=0.0)
    ap.add_argument("--capacity-factor", type=float, default=0.0,
                    help="receive-buffer rows / local (token, expert) pairs; 0 = auto (retry with a larger one on overflow)")
    ap.add_argument("--shadow-experts", type=int, default=8,
                    help="max hot experts per layer and step that are processed data-parallel on every rank (0 = static placement)")
    ap.add_argument("--shadow-tol", type=float, default=1.1,
                    help="shadow selection stops once the most loaded rank is within this factor of the mean load")
    ap.add_argument("--expert-dtype", choices=["bf16", "fp8"], default="bf16",
                    help="fp8 = forward expert GEMMs on block-scaled FP8 tensor cores (MXFP8)")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons of this rank's GPU while the timed region runs"""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            loaded = sorted(sm)[len(sm) // 2]
            out = dict(sm_mhz=loaded, sm_max_mhz=max(mx), power_w_max=max(power) if power else None,
                       samples=len(sm), reasons=sorted(reasons))
        return out


def dist_setup(n_gpus):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
        try:  # run next to the GPU: the pinned staging buffers of the end-to-end path are placed by first touch
            import lah_b200  # noqa
            from lah_b200.utils.affinity import bind_to_gpu_numa_node
            bind_to_gpu_numa_node(local_rank)
        except Exception:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo",
                                device_id=torch.device("cuda", local_rank) if torch.cuda.is_available() else None)
    return rank, world, local_rank


def max_over_ranks(value, world):
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier_sync(world):
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn_step, steps, world):
    """EXACTLY `steps` calls bracketed by barrier + synchronize; CUDA events on the launching stream; max over ranks"""
    barrier_sync(world)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        fn_step(i)
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end)
    barrier_sync(world)
    return max_over_ranks(ms, world)


# ------------------------------------------------------------------------------------------------ our engine
def run_ours(args):
    rank, world, local_rank = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "ours", "unavailable": "no CUDA device visible"}))
        return
    import lah_b200  # noqa
    from lah_b200.ops import native
    from lah_b200.parallel.engine import DMoEConfig
    from lah_b200.parallel.trainer import DMoETrainer

    B = args.batch_per_gpu
    # Expert load is imbalanced by nature (deep layers route most tokens through a few hot experts), so the rank hosting a
    # hot expert receives far more than its share of rows.  Nothing is ever dropped: if a receive buffer overflows the
    # engine raises, and we re-run the WHOLE measurement with bigger buffers.
    if args.capacity_factor > 0:
        factors = [args.capacity_factor]
    elif os.environ.get("LAH_BENCH_FACTORS"):
        factors = [float(f) for f in os.environ["LAH_BENCH_FACTORS"].split(",")]
    elif args.shadow_experts > 0 and world > 1:   # balanced by shadowing: every rank receives ~ its own share
        factors = sorted({1.5, min(2.5, float(world)), float(world)})
    else:
        factors = sorted({min(3.0, float(world)), min(5.0, float(world)), float(world)})
    for attempt, factor in enumerate(factors):
        try:
            _measure_ours(args, rank, world, local_rank, B, factor)
            break
        except RuntimeError as e:
            if "overflow" not in str(e) or attempt == len(factors) - 1:
                raise
            if rank == 0:
                print(f"[bench] receive buffers overflowed at capacity_factor={factor}; retrying with {factors[attempt + 1]}",
                      file=sys.stderr)
            import gc
            gc.collect()
            torch.cuda.empty_cache()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def _measure_ours(args, rank, world, local_rank, B, capacity_factor):
    import lah_b200  # noqa
    from lah_b200.ops import native
    from lah_b200.parallel.engine import DMoEConfig
    from lah_b200.parallel.trainer import DMoETrainer
    cfg = DMoEConfig(hidden=args.hidden, grid_size=tuple(args.grid), k=args.k, num_layers=args.layers,
                     tokens_per_rank=B, capacity_factor=capacity_factor, failure_rate=args.failure_rate,
                     gate_mode=args.gate, shadow_experts=args.shadow_experts, shadow_tol=args.shadow_tol,
                     expert_dtype=args.expert_dtype)
    trainer = DMoETrainer(cfg)
    gen = torch.Generator().manual_seed(1234 + rank)
    n_batches = 4
    xs_host = [torch.randn(B, cfg.in_features, generator=gen).pin_memory() for _ in range(n_batches)]
    ys_host = [torch.randint(0, cfg.num_classes, (B,), generator=gen).pin_memory() for _ in range(n_batches)]
    xs_dev = [x.cuda(non_blocking=True) for x in xs_host]
    ys_dev = [y.cuda(non_blocking=True) for y in ys_host]

    def step_device(i):
        trainer.train_step_device(xs_dev[i % n_batches], ys_dev[i % n_batches])

    losses = []

    pending = []

    def step_e2e(i):
        # end to end through the public API: pinned host inputs -> H2D (the next step's inputs cross PCIe on a copy stream
        # while this step computes) -> step -> D2H loss.  The loss of step i is read (blocking) right after step i+1 has
        # been enqueued, so every step's result reaches the host inside the timed region but the host stays one step ahead
        nxt = (i + 1) % n_batches
        pending.append(trainer.train_step_async(xs_host[i % n_batches], ys_host[i % n_batches],
                                                prefetch=(xs_host[nxt], ys_host[nxt])))
        if len(pending) > 1:
            losses.append(pending.pop(0).result())

    def drain_e2e():
        while pending:
            losses.append(pending.pop(0).result())

    def check_all_ranks():
        code = int(trainer.ctx.status.item())
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([code & 1, code & 2], device="cuda")  # NCCL has no bitwise-or: MAX per status bit
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            code = int(t[0].item()) | int(t[1].item())
        if code:
            trainer.ctx.status.fill_(code)
            try:
                trainer.ctx.check_status()
            finally:
                trainer.ctx.heap.close()

    barrier_sync(world)   # ranks finish building their host batches at different times: start the first step together
    for i in range(args.warmup):
        step_device(i)
    check_all_ranks()
    sampler = ClockSampler(local_rank)
    sampler.start()
    native.reset_launches()
    trainer.ctx.exposed_wait_ms(reset=True)
    profiling = bool(os.environ.get("LAH_CUDA_PROFILE"))  # ncu --profile-from-start off
    if profiling:
        torch.cuda.profiler.start()
    ms = timed(step_device, args.steps, world)
    if profiling:
        torch.cuda.profiler.stop()
    launches = native.launches()
    exposed_ms = max_over_ranks(trainer.ctx.exposed_wait_ms(reset=True) / args.steps, world)
    clocks = sampler.stop()
    check_all_ranks()
    global_batch = B * world
    value = global_batch * args.steps / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        for i in range(2):
            step_e2e(i)
        drain_e2e()

        def e2e_steps(i):   # the last timed step also waits for its own loss: all K results are on the host at the end
            step_e2e(i)
            if i == args.steps - 1:
                drain_e2e()

        ms_e2e = timed(e2e_steps, args.steps, world)
        check_all_ranks()
        e2e = {"value": global_batch * args.steps / (ms_e2e / 1e3), "unit": "samples/s",
               "ms_per_step": ms_e2e / args.steps,
               "h2d_bytes_per_step": int(xs_host[0].numel() * 4 + ys_host[0].numel() * 8),
               "d2h_bytes_per_step": 4, "last_loss": losses[-1] if losses else None}

    # per-stage breakdown of ONE extra (untimed) step: CUDA events between the stages, synchronised afterwards
    trainer.ctx.timer.enabled = True
    step_device(0)
    stage_ms = {k: round(v, 3) for k, v in trainer.last_stage_ms.items()}
    trainer.ctx.timer.enabled = False
    routing = []
    for block in trainer.model.blocks:  # tokens-per-expert histogram of the last step (observability)
        rows = block.ws.step_rows.float()
        routing.append({"active_experts": int((rows > 0).sum()), "max_rows": int(rows.max()), "mean_rows": float(rows.mean()),
                        "padded_rows": int(block.ws.total_rows.item()),
                        "shadowed_experts": int((block.ws.shadow_info.view(-1, 4)[:, 0] >= 0).sum())})
    if rank == 0:
        out = {
            "metric": "DMoE training samples/sec (whole job, device-timed, max over ranks)",
            "impl": "ours", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / 16.8, "dtype": "bf16", "data": "synthetic (MNIST-shaped 784-d fp32 rows, random-init weights)",
            "config": {"model": f"Linear(784,{cfg.hidden}) -> {cfg.num_layers} x DMoE[{cfg.num_experts} experts "
                                f"FeedforwardBlock({cfg.hidden}), top-{cfg.k}] -> LayerNorm -> Linear({cfg.hidden},10); "
                                "fwd+bwd+per-expert AMSGrad+trainer AMSGrad",
                       "global_batch": global_batch, "seq_len": 1, "parallelism": f"ep{world}+dp{world}",
                       "capacity_factor": capacity_factor, "shadow_experts": cfg.shadow_experts if world > 1 else 0, "shadow_tol": cfg.shadow_tol,
                       "expert_gemm_dtype": cfg.expert_dtype + (" forward, bf16 dgrad/wgrad" if cfg.expert_dtype == "fp8" else ""),
                       "experts_total": cfg.num_experts * cfg.num_layers, "failure_rate": cfg.failure_rate,
                       "gate": cfg.gate_mode + (" (LayerNorm(x) @ normalize(keys); gate params not trained, exactly like the reference's EmulatedDMoE)" if cfg.gate_mode == "emulator" else " (trainable product-key proj, lib.GatingFunction)"),
                       "l2_policy": "working set per step (>35 GB of expert state + >10 GB activations) exceeds the 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "routing_rank0": routing,
            "exposed_comm_wait_ms_per_step": exposed_ms, "stage_ms_rank0": stage_ms,
            "baseline_note": "vs_baseline = value / 16.8 samples/s (reference notebook dmoe64x4, BASELINE.md)",
        }
        print(json.dumps(out))
    return True


# ------------------------------------------------------------------------------------------------ NCCL + cuBLAS baseline
def run_baseline(args):
    """same model / step through parallel/baseline.py: product-key gate, index_select permute, all_to_all_single (NCCL),
    per-expert F.linear under bf16 autocast (cuBLAS), one torch Adam(amsgrad) per expert"""
    rank, world, local_rank = dist_setup(args.gpus)
    cuda = torch.cuda.is_available()
    if not cuda and not os.environ.get("LAH_BENCH_ALLOW_CPU"):
        print(json.dumps({"impl": "baseline", "unavailable": "no CUDA device visible"}))
        return
    import lah_b200  # noqa
    from lah_b200.parallel.engine import DMoEConfig
    from lah_b200.parallel.baseline import BaselineTrainer
    B = args.batch_per_gpu
    grid = tuple(args.grid) if len(args.grid) > 1 else (8, 8) if args.grid == [64] else tuple(args.grid)
    cfg = DMoEConfig(hidden=args.hidden, grid_size=grid, k=args.k, num_layers=args.layers, tokens_per_rank=B,
                     failure_rate=args.failure_rate, gate_mode="product_key")
    trainer = BaselineTrainer(cfg, dtype=torch.bfloat16 if cuda else torch.float32)
    gen = torch.Generator().manual_seed(1234 + rank)
    dev = trainer.device
    xs = [torch.randn(B, cfg.in_features, generator=gen).to(dev) for _ in range(2)]
    ys = [torch.randint(0, cfg.num_classes, (B,), generator=gen).to(dev) for _ in range(2)]

    def step(i):
        trainer.train_step_device(xs[i % 2], ys[i % 2])

    for i in range(args.warmup):
        step(i)
    if cuda:
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms = timed(step, args.steps, world)
        clocks = sampler.stop()
    else:  # CPU smoke path of this arm (tests)
        t0 = time.time()
        for i in range(args.steps):
            step(i)
        ms, clocks = (time.time() - t0) * 1e3, {}
    value = B * world * args.steps / (ms / 1e3)
    if rank == 0:
        print(json.dumps({
            "metric": "DMoE training samples/sec (whole job, device-timed, max over ranks)", "impl": "baseline",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": value / 16.8,
            "dtype": "bf16 autocast" if cuda else "fp32", "data": "synthetic (MNIST-shaped 784-d fp32 rows, random-init weights)",
            "config": {"model": f"same model through torch.topk + all_to_all_single + cuBLAS + torch Adam; grid {grid}",
                       "global_batch": B * world, "seq_len": 1, "parallelism": f"ep{world}+dp{world} (NCCL all_to_all_single)"},
            "clocks": clocks, "gpu_launches": 0}))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "experiments", "convergence")):
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref missing (run baseline/install_reference.sh)"}))
        return
    rank, world, local_rank = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        print(json.dumps({"impl": "reference", "unavailable": "no CUDA device visible"}))
        return
    sys.path.insert(0, ref_root)
    from baseline.ref_bench import build_reference_trainer
    B = args.ref_batch
    step, info = build_reference_trainer(hidden=args.hidden, num_experts=1 if not args.grid else int(torch.tensor(args.grid).prod()),
                                         num_active=args.k, num_layers=args.layers, batch_size=B,
                                         failure_rate=args.failure_rate, seed=1337 + rank)
    for i in range(args.warmup):
        step(i, e2e=False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(lambda i: step(i, e2e=False), args.steps, world)
    clocks = sampler.stop()
    value = B * world * args.steps / (ms / 1e3)
    e2e = None
    if not args.no_e2e:
        ms_e2e = timed(lambda i: step(i, e2e=True), args.steps, world)
        e2e = {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e / args.steps,
               "h2d_bytes_per_step": B * 784 * 4 + B * 8, "d2h_bytes_per_step": 4}
    if rank == 0:
        print(json.dumps({
            "metric": "DMoE training samples/sec (whole job, device-timed, max over ranks)", "impl": "reference",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": value / 16.8,
            "dtype": "fp32", "data": "synthetic (MNIST-shaped 784-d fp32 rows, random-init weights)",
            "config": {"model": info, "global_batch": B * world, "seq_len": 1,
                       "parallelism": "1 independent replica per GPU (the reference emulator is single-process)"},
            "clocks": clocks, "gpu_launches": 0, "e2e": e2e}))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "baseline":
        run_baseline(a)
    elif a.impl == "reference":
        try:
            run_reference(a)
        except Exception as e:  # the reference arm must never fail the driver
            print(json.dumps({"impl": "reference", "unavailable": f"{type(e).__name__}: {e}"[:300]}))
    else:
        run_ours(a)
