"""torchrun -n 2: kill one rank's participation mid-run; the other detects it (bounded flag waits + device-resident heartbeat
table), excludes it and keeps training on its own experts.  Prints FAULT_OK on rank 0.

Timeline: steps 0..4 both ranks | rank 1 stops stepping and heart-beating | rank 0: step 5 times out after peer_timeout_ms
(no optimizer update is applied from the partial step) -> recover() -> steps 6..17 alone, loss keeps falling."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import lah_b200  # noqa
from lah_b200.parallel import engine as E
from lah_b200.parallel.trainer import DMoETrainer


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    cfg = E.DMoEConfig(hidden=512, grid_size=(16,), k=4, num_layers=2, tokens_per_rank=256, gate_mode="emulator",
                       capacity_factor=float(world), peer_timeout_ms=300)
    tr = DMoETrainer(cfg, use_graph=("--graph" in sys.argv))
    gen = torch.Generator().manual_seed(4242)
    protos = torch.randn(10, cfg.in_features, generator=gen)
    gen = torch.Generator().manual_seed(rank)

    def batch():
        y = torch.randint(0, 10, (256,), generator=gen)
        return (protos[y] + 3.0 * torch.randn(256, cfg.in_features, generator=gen)).cuda(), y.cuda()

    tr.ctx.heartbeat()
    losses = []
    for step in range(5):
        tr.ctx.heartbeat()
        losses.append(float(tr.train_step_device(*batch())))
    torch.cuda.synchronize()
    assert not tr.step_failed(), "healthy steps must not time out"
    dist.barrier()
    if rank != 0:
        # "crash": no more steps, no more heartbeats; the process only waits for the end of the test
        dist.barrier()
        tr.close()
        dist.destroy_process_group()
        return
    # ---- rank 0 carries on
    p_before = tr.model.blocks[0].shard.p.clone()
    t0 = time.time()
    tr.ctx.heartbeat()
    tr.train_step_device(*batch())
    torch.cuda.synchronize()
    stall = time.time() - t0
    failed = tr.step_failed()
    untouched = bool(torch.equal(p_before, tr.model.blocks[0].shard.p))   # the failed step applied no expert update
    time.sleep(0.7)
    dead = tr.recover(max_age=0.5)
    after = []
    for step in range(12):
        tr.ctx.heartbeat()
        after.append(float(tr.train_step_device(*batch())))
    torch.cuda.synchronize()
    ok_after = not tr.step_failed()
    alive = int(tr.ctx.alive.sum())
    print(dict(before=[round(v, 3) for v in losses], failed_step_detected=failed, stall_s=round(stall, 2),
               no_update_from_failed_step=untouched, excluded=dead, alive_experts=alive,
               after=[round(v, 3) for v in after], ok_after=ok_after), flush=True)
    good = failed and untouched and dead == [1] and ok_after and alive == cfg.num_experts // world and after[-1] < after[0] \
        and stall < 10.0
    print("FAULT_OK" if good else "FAULT_FAILED", flush=True)
    dist.barrier()
    tr.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
