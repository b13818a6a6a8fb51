"""Diagnostics for the multi-GPU bench legs: all-gather time, concurrent H2D/D2H bandwidth."""
import os, time, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
n=260046720//8
x=torch.randn(n,dtype=torch.float64,device=dev); g=torch.empty((world,n),dtype=torch.float64,device=dev)
for _ in range(3): dist.all_gather_into_tensor(g,x)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): dist.all_gather_into_tensor(g,x)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/5
print(f"rank {rank}: all_gather {n*8/1e6:.0f} MB/rank: {ms:.3f} ms -> {n*8*(world-1)/ms/1e6:.1f} GB/s in", flush=True)
h=torch.empty(n,dtype=torch.float64).pin_memory()
for name,fn in (("d2h",lambda: h.copy_(x,non_blocking=True)),("h2d",lambda: x.copy_(h,non_blocking=True))):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(f"rank {rank}: {name} {n*8/1e6:.0f} MB: {ms:.3f} ms -> {n*8/ms/1e6:.1f} GB/s (all ranks concurrently)", flush=True)
dist.destroy_process_group()
