"""Does symmetric memory on this box come with an NVLS multicast mapping? (torchrun, >= 2 ranks)"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
t = symm_mem.empty((1 << 20,), dtype=torch.uint8, device="cuda")
h = symm_mem.rendezvous(t, dist.group.WORLD)
mc = getattr(h, "multicast_ptr", None)
print(f"rank {rank}: world {h.world_size} multicast_ptr {mc} has_multicast_support "
      f"{getattr(symm_mem, 'has_multicast_support', lambda *a: 'n/a')('cuda', torch.cuda.current_device()) if hasattr(symm_mem, 'has_multicast_support') else 'n/a'}", flush=True)
dist.barrier(); dist.destroy_process_group()
