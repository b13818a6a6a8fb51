"""Write-out gather of the finished-track blocks across the GPUs of one node (reference: util/compute.py:233-242 collects
one 9-tuple per dask worker and concatenates them on the host).

Storms are independent, so this is the ONLY exchange of the hot path.  Two transports, same result:

  * PeerGather   -- every rank writes its block straight into every peer's (or only rank 0's) gather buffer with
                    device-to-device copies over NVLink/NVSwitch.  The buffers are torch allocations shared between the
                    processes through CUDA IPC, the copies run on the COPY ENGINES of a side stream: no SM is needed, so
                    the exchange overlaps the persistent integrator of the next batch, which owns every SM's register
                    file (an NCCL all-gather's CTAs cannot become resident beside it: measured in round 1,
                    profiles/r01_n2_gather_modes.txt).
  * nccl_gather  -- `dist.all_gather_into_tensor`, the plain collective (the baseline the above replaces).

torch is used for device memory, streams, IPC handles and the process group: plumbing.
"""
import os

import numpy as np


def bind_to_gpu_numa(local_rank):
    """Pin this process (and so its pinned host allocations, first-touch) to the CPUs of the NUMA node its GPU hangs
    off.  Eight ranks staging their results through one node's memory was the round-1 end-to-end bottleneck.
    Returns a short description of what was done (for the bench record)."""
    try:
        import torch
        props = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa node unknown for %s" % bus
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return "numa node %d has no allowed cpu" % node
        os.sched_setaffinity(0, allowed)
        return "gpu %s -> numa node %d, %d cpus" % (bus, node, len(allowed))
    except Exception as e:                                     # no sysfs / no permission: run unbound
        return "unbound (%s)" % type(e).__name__


_TORCH_DTYPES = {0: ("|u1", "uint8"), 1: ("<i4", "int32"), 2: ("<i8", "int64")}


class _DevView:
    """A device buffer owned by libtcrisk.so, exposed to torch without a copy (__cuda_array_interface__)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = dict(shape=(int(count),), typestr=typestr, data=(int(ptr), False), version=2)


def device_tensor(ptr, count, dtype_code, device):
    import torch
    typestr, name = _TORCH_DTYPES[dtype_code]
    return torch.as_tensor(_DevView(ptr, count, typestr), device=device)


def dist_allreduce(device, group=None):
    """The all-reduce callback of Engine.set_shard over torch.distributed (NCCL on GPUs): in place, stream-ordered on
    torch's current stream -- which must be the stream the engine runs on (Engine.set_stream)."""
    import torch
    import torch.distributed as dist

    def fn(ptr, count, dtype_code, op, stream):
        t = device_tensor(ptr, count, dtype_code, device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN if op == 1 else dist.ReduceOp.SUM, group=group)
    return fn


def merge_sharded_block(block, group=None):
    """Blocks of within-year-sharded ranks -> the complete block on every rank: rows a rank did not integrate are
    all-zero bits, so the merge is an INTEGER sum of the float64 bit patterns (exact for NaN padding and signed zeros)."""
    import torch
    import torch.distributed as dist
    dist.all_reduce(block.view(torch.int64), op=dist.ReduceOp.SUM, group=group)
    return block


def nccl_gather(block, out=None):
    """all_gather_into_tensor of equal-sized 1-D device blocks -> [world][n] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world, block.numel()), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, block)
    return out


class PeerGather:
    """Gather buffers of `depth` generations, [world][n] on every destination rank, written by peers through CUDA IPC.

    push(block, gen)   enqueue, on this rank's copy stream (after the current stream's work), one device-to-device copy of
                       `block` into slot [rank] of every destination's buffer of generation gen % depth
    wait_local(gen)    current stream waits until this rank's pushes of that generation have left `block` (the block --
                       one per generation -- may then be overwritten)
    finish()           all ranks' pushes of all generations have landed everywhere (stream sync + barrier)
    buffer(gen)        this rank's [world][n] gather buffer (destinations only)

    dst = "all" (all-gather) or "root" (gather to rank 0: what a write-out by rank 0 needs)."""

    def __init__(self, n, dtype, device, depth=2, dst="all", dst_depth=None):
        import torch
        import torch.distributed as dist
        from torch.multiprocessing.reductions import rebuild_cuda_tensor, reduce_tensor
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.n, self.depth = int(n), depth
        self.dst_depth = depth if dst_depth is None else dst_depth       # generations kept at the destination
        self.dst_ranks = list(range(self.world)) if dst == "all" else [0]
        self.mine = None
        handle = None
        if self.rank in self.dst_ranks:
            self.mine = torch.empty((self.dst_depth, self.world, self.n), dtype=dtype, device=device)
            fn, args = reduce_tensor(self.mine)
            assert fn is rebuild_cuda_tensor
            handle = args
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)
        self.peers = {}
        for r in self.dst_ranks:
            if r == self.rank:
                self.peers[r] = self.mine
            else:
                args = list(handles[r])
                args[6] = device.index if hasattr(device, "index") else int(device)     # storage_device: map into MY context
                self.peers[r] = rebuild_cuda_tensor(*args)
        self.stream = torch.cuda.Stream(device=device)
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.pushed = [False] * depth
        self.bytes_pushed = 0

    def push(self, block, gen):
        torch = self.torch
        self.stream.wait_stream(torch.cuda.current_stream())
        g = gen % self.depth
        with torch.cuda.stream(self.stream):
            # start with the right-hand neighbour so that the ranks do not all write to the same peer at once
            for k in range(len(self.dst_ranks)):
                r = self.dst_ranks[(self.rank + 1 + k) % len(self.dst_ranks)]
                self.peers[r][gen % self.dst_depth, self.rank].copy_(block, non_blocking=True)
            self.done[g].record(self.stream)
        self.pushed[g] = True
        self.bytes_pushed += block.numel() * block.element_size() * len(self.dst_ranks)

    def wait_local(self, gen):
        if self.pushed[gen % self.depth]:
            self.torch.cuda.current_stream().wait_event(self.done[gen % self.depth])

    def finish(self):
        self.stream.synchronize()
        self.dist.barrier()

    def buffer(self, gen):
        return None if self.mine is None else self.mine[gen % self.dst_depth]
