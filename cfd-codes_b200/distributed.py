"""Host-side plumbing for multi-GPU runs: one process per GPU (torchrun), `torch.distributed` for the
rendezvous only.  The data path (halo exchange of the five face-crossing populations per slab face,
residual all-reduce) is NCCL inside liblbm_b200.so; nothing here touches the lattice.

The reference's only distributed program is the 2D one (MPI row slabs, 2D main.cpp:97-125, 208-270);
this is the same decomposition along the slowest axis of the 3D lattice (z, LU4 main.c:48)."""
import os

import numpy as np

from . import NCCL_ID_BYTES, PEER_BLOB_BYTES, nccl_unique_id, slab_range


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs that are local to its GPU (NVML's ideal affinity), so that the pinned
    host buffers it allocates next live on the GPU's own NUMA node and PCIe transfers do not cross the
    socket interconnect.  Returns the previous affinity set (pass it to os.sched_setaffinity to undo), or
    None if nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        old = os.sched_getaffinity(0)
        new = cpus & old
        if not new or new == old:
            return None
        os.sched_setaffinity(0, new)
        return old
    except Exception:
        return None


def init_process_group(backend=None):
    """Rendezvous over 127.0.0.1 (the container hostname may not resolve)."""
    import torch
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if not dist.is_initialized():
        if backend is None:
            backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend)
    return dist


def broadcast_nccl_id(make_id=nccl_unique_id):
    """Rank 0 creates the ncclUniqueId (through the C ABI), everybody receives the 128 bytes."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(NCCL_ID_BYTES, dtype=torch.uint8)
    if dist.get_rank() == 0:
        raw = make_id()
        assert len(raw) == NCCL_ID_BYTES
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.tolist())


def gather_slabs(local, n_slowest, plane_elems, nfields):
    """Gather compact per-rank slabs [nfields][k1-k0][plane] onto rank 0 as [nfields][n][plane]
    (the reference's LU4 / LU3 order).  Returns the full array on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    k0, k1 = slab_range(n_slowest, rank, world)
    local = np.ascontiguousarray(local, dtype=np.float64).reshape(nfields, k1 - k0, plane_elems)
    if rank == 0:
        full = np.empty((nfields, n_slowest, plane_elems), dtype=np.float64)
        full[:, k0:k1] = local
        for r in range(1, world):
            a, b = slab_range(n_slowest, r, world)
            buf = torch.empty(nfields * (b - a) * plane_elems, dtype=torch.float64)
            dist.recv(buf, src=r)
            full[:, a:b] = buf.numpy().reshape(nfields, b - a, plane_elems)
        return full
    dist.send(torch.from_numpy(local.reshape(-1)), dst=0)
    return None


def allreduce_max(x):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def connect_direct_halo(sim):
    """Switch a multi-GPU Cavity3D to the direct NVLink halo: all-gather the per-rank export blobs
    (torch.distributed, CPU tensors) and connect each rank to its two z-neighbours."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = torch.frombuffer(bytearray(sim.peer_export()), dtype=torch.uint8)
    blobs = [torch.zeros(PEER_BLOB_BYTES, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(blobs, mine)
    below = bytes(blobs[rank - 1].tolist()) if rank > 0 else None
    above = bytes(blobs[rank + 1].tolist()) if rank + 1 < world else None
    dist.barrier()
    sim.peer_connect(below, above)
    dist.barrier()


def connect_direct_halo_via(sim, dist, device):
    """Same as connect_direct_halo for a process group that only has the NCCL backend (bench.py):
    the blobs are all-gathered as CUDA byte tensors."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = torch.tensor(list(sim.peer_export()), dtype=torch.uint8, device=device)
    blobs = [torch.zeros(PEER_BLOB_BYTES, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(blobs, mine)
    below = bytes(blobs[rank - 1].cpu().tolist()) if rank > 0 else None
    above = bytes(blobs[rank + 1].cpu().tolist()) if rank + 1 < world else None
    dist.barrier()
    sim.peer_connect(below, above)
    dist.barrier()
