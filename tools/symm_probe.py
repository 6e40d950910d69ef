"""torchrun -n W: does torch symmetric memory (VMM + multicast) rendezvous work on this box? prints pointers."""
import os, sys, json
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
out = dict(rank=rank, world=world)
try:
    t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", lr))
    h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    out.update(buffer_ptrs=[hex(p) for p in h.buffer_ptrs], multicast_ptr=hex(h.multicast_ptr or 0),
               signal_pad_ptrs=[hex(p) for p in h.signal_pad_ptrs], buffer_size=h.buffer_size,
               has_mc=bool(h.multicast_ptr))
    big = symm_mem.empty(4 << 30, dtype=torch.uint8, device=torch.device("cuda", lr))
    hb = symm_mem.rendezvous(big, dist.group.WORLD.group_name)
    out.update(big_ok=True, big_mc=hex(hb.multicast_ptr or 0))
except Exception as e:
    out.update(error=f"{type(e).__name__}: {e}"[:500])
print(json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
