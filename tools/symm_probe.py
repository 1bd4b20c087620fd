"""torchrun probe: does torch's symmetric memory give peer-mapped device pointers on this box?"""
import os
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

t = symm_mem.empty(1 << 20, dtype=torch.uint8, device="cuda:%d" % local)
t.zero_()
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name if hasattr(dist.group.WORLD, "group_name") else dist.group.WORLD)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs][:2],
      "rank", hdl.rank, "world", hdl.world_size, flush=True)
# write into the peer's buffer through its mapped pointer
peer = (rank + 1) % world
buf = hdl.get_buffer(peer, (16,), torch.float32, 0)
buf.fill_(float(rank + 1))
hdl.barrier()
dist.barrier()
torch.cuda.synchronize()
mine = t[:64].view(torch.float32)
print(rank, "sees", mine[:4].tolist(), "can_access_peer", torch.cuda.can_device_access_peer(local, peer), flush=True)
dist.destroy_process_group()
