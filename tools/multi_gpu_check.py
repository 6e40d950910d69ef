"""torchrun -n W: the fused P2P engine on W GPUs must reproduce the single-GPU engine on the concatenated batch.

Every rank feeds a different slice of one global batch; rank r hosts experts [r*E/W, (r+1)*E/W).  We compare, for one
DMoE layer: outputs y, input gradients dx, gate gradients, and the expert parameters after the optimizer step, against a
single-process run of the SAME layer code over the whole batch (world=1 path, all experts local)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import lah_b200  # noqa
from lah_b200.parallel import engine as E


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    B = 256
    force_shadow = "--force-shadow" in sys.argv  # shadow as many experts as there are slots (exercises the replica path)
    small = "--small" in sys.argv   # weight-streaming expert path (swap-AB GEMMs + fused wgrad/AMSGrad; no shadowing)
    cfg = E.DMoEConfig(hidden=512, grid_size=(4, 4), k=4, num_layers=1, tokens_per_rank=B, capacity_factor=float(max(4, world)),
                       shadow_experts=4, shadow_tol=0.0 if force_shadow else 1.1, shadow_min_rows=1 if force_shadow else 64,
                       expert_path="small" if small else "big")
    ctx = E.EngineContext(cfg)
    torch.manual_seed(0)  # identical gate on every rank (DMoETrainer does the same)
    layer = E.FusedDMoE(cfg, ctx).cuda()
    gen = torch.Generator().manual_seed(0)
    x_all = torch.randn(world * B, 512, generator=gen).to(torch.bfloat16)
    g_all = torch.randn(world * B, 512, generator=gen).to(torch.bfloat16)
    x = x_all[rank * B: (rank + 1) * B].cuda().requires_grad_(True)
    y = layer(x)
    y.backward(g_all[rank * B: (rank + 1) * B].cuda())
    torch.cuda.synchronize()
    ctx.check_status()
    shadowed = int((layer.ws.shadow_info.view(-1, 4)[:, 0] >= 0).sum())
    # the kernel's hot-expert selection must equal the host model (parallel/balance.py) on the exchanged count table
    from lah_b200.parallel.balance import shadow_plan
    counts = ctx.cnt_all[:world].cpu().tolist()
    plan, _ = shadow_plan(counts, ctx.E_loc, ctx.S, tol=cfg.shadow_tol, min_rows=cfg.shadow_min_rows)
    got = [int(e) for e in layer.ws.shadow_info.view(-1, 4)[:, 0].cpu().tolist() if e >= 0]
    plan_ok = plan == got or small
    # gather what the distributed run produced
    ys = [torch.empty_like(y) for _ in range(world)]
    dxs = [torch.empty_like(x.grad) for _ in range(world)]
    dist.all_gather(ys, y.detach().contiguous())
    dist.all_gather(dxs, x.grad.contiguous())
    gw = layer.proj.weight.grad.clone()
    dist.all_reduce(gw)
    w1 = layer.shard.views["w1"][:layer.E_loc].clone()
    w1s = [torch.empty_like(w1) for _ in range(world)]
    dist.all_gather(w1s, w1)
    b2 = layer.shard.views["b2"][:layer.E_loc].clone()
    b2s = [torch.empty_like(b2) for _ in range(world)]
    dist.all_gather(b2s, b2)
    steps = [torch.empty_like(layer.shard.step) for _ in range(world)]
    dist.all_gather(steps, layer.shard.step)
    ok = True
    if rank == 0:
        # single-GPU reference in the same process: a fresh world-1 context is impossible inside an initialised group,
        # so use the PyTorch oracle of the layer (all experts local) on the whole batch
        ref_cfg = cfg
        ref = E.FusedDMoE(ref_cfg, None, device=torch.device("cuda")).cuda()
        ref.proj.load_state_dict(layer.proj.state_dict())
        ref.train()
        xr = x_all.cuda().float().requires_grad_(True)
        yr = ref(xr)
        yr.backward(g_all.cuda().float())
        ref.apply_expert_gradients_ref()
        errs = dict(y=rel(torch.cat(ys), yr), dx=rel(torch.cat(dxs), xr.grad), dproj=rel(gw, ref.proj.weight.grad),
                    w1_mean_abs=(torch.cat(w1s) - ref.shard.views["w1"]).abs().mean().item(),
                    b2_max_abs=(torch.cat(b2s) - ref.shard.views["b2"]).abs().max().item(),
                    steps=bool((torch.cat(steps).cpu() == ref.shard.step.cpu()).all()))
        ok = errs["y"] < 2e-2 and errs["dx"] < 3e-2 and errs["dproj"] < 5e-2 and errs["w1_mean_abs"] < 1e-4 and errs["b2_max_abs"] < 2.5e-3 and errs["steps"]
        ok = ok and (shadowed > 0 or not force_shadow) and plan_ok
        print("multi_gpu_check", dict(path="small" if small else "big", force_shadow=force_shadow, shadowed_experts=shadowed, plan_matches_host_model=plan_ok,
                                      plan=plan, kernel=got), errs, flush=True)
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
