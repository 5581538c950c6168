import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm.empty(1<<20, dtype=torch.uint8, device=torch.device("cuda", local))
h = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "multicast_ptr", hex(int(h.multicast_ptr)), "buffer_ptrs", [hex(int(p)) for p in h.buffer_ptrs][:2], flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
