"""Multi-GPU plumbing: reads shard by rank, the index is replicated per GPU, and the only
collective is the final gather of per-rank SMEM records (north_star (4)).

One process per GPU (torchrun); torch.distributed is the transport (NCCL on GPUs, gloo in the CPU
tests).  There is no data-path collective: every rank searches its own contiguous block of reads.
"""
import numpy as np
import torch
import torch.distributed as dist

from .engine import RECORD_DTYPE


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) so that the pinned host buffers of the
    end-to-end path are first-touched on the GPU's NUMA node.  Best effort: returns the CPU set, or None."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def shard_range(n_reads, rank, world):
    """Contiguous block [lo, hi) of reads for `rank`: rank r gets reads [r*N/G, (r+1)*N/G)."""
    return (n_reads * rank) // world, (n_reads * (rank + 1)) // world


def gather_records(records, counts_per_read, dst=0, group=None, device=None):
    """Gather variable-length record arrays to `dst`.

    records: structured array (RECORD_DTYPE) of this rank, read ids already global (read_id_base);
    counts_per_read: int64 array, records per local read.  Returns (records, counts) concatenated
    in rank order on dst, (None, None) elsewhere.  Two collectives: all_gather of the two sizes,
    then gather of the payloads padded to the largest shard.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    rec_u8 = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).copy()) if not torch.is_tensor(records) else records
    cnt = torch.from_numpy(np.ascontiguousarray(counts_per_read, dtype=np.int64)) if not torch.is_tensor(counts_per_read) else counts_per_read
    rec_u8, cnt = rec_u8.to(device), cnt.to(device)
    sizes = torch.tensor([rec_u8.numel(), cnt.numel()], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    max_rec, max_cnt = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    pad_rec = torch.empty(max(max_rec, 1), dtype=torch.uint8, device=device)
    pad_rec[: rec_u8.numel()] = rec_u8
    pad_cnt = torch.empty(max(max_cnt, 1), dtype=torch.int64, device=device)
    pad_cnt[: cnt.numel()] = cnt
    recv_rec = [torch.empty_like(pad_rec) for _ in range(world)] if rank == dst else None
    recv_cnt = [torch.empty_like(pad_cnt) for _ in range(world)] if rank == dst else None
    dist.gather(pad_rec, recv_rec, dst=dst, group=group)
    dist.gather(pad_cnt, recv_cnt, dst=dst, group=group)
    if rank != dst:
        return None, None
    n_rec = [int(all_sizes[r, 0]) for r in range(world)]
    n_cnt = [int(all_sizes[r, 1]) for r in range(world)]
    if device.type == "cuda":
        # one pinned destination per array, every shard copied straight to its final place (no pageable staging, no concatenate)
        out_rec = torch.empty(max(sum(n_rec), 1), dtype=torch.uint8, pin_memory=True)
        out_cnt = torch.empty(max(sum(n_cnt), 1), dtype=torch.int64, pin_memory=True)
        a = b = 0
        for r in range(world):
            out_rec[a:a + n_rec[r]].copy_(recv_rec[r][: n_rec[r]], non_blocking=True)
            out_cnt[b:b + n_cnt[r]].copy_(recv_cnt[r][: n_cnt[r]], non_blocking=True)
            a += n_rec[r]
            b += n_cnt[r]
        torch.cuda.synchronize(device)
        return out_rec.numpy()[: sum(n_rec)].view(RECORD_DTYPE), out_cnt.numpy()[: sum(n_cnt)]
    recs = np.concatenate([recv_rec[r][: n_rec[r]].numpy() for r in range(world)]).view(RECORD_DTYPE)
    cnts = np.concatenate([recv_cnt[r][: n_cnt[r]].numpy() for r in range(world)])
    return recs, cnts
