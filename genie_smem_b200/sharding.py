"""Multi-GPU plumbing: reads shard by rank, the index is replicated per GPU, and the only
collective is the final gather of per-rank SMEM records (north_star (4)).

One process per GPU (torchrun).  torch.distributed carries the bootstrap (the 128-byte NCCL id, the
CUDA IPC handle of the destination buffer) and is the transport of the gloo CPU tests; on GPUs the
data path is the library's own: gsm_gather_records (exact-size ncclSend / ncclRecv into a
preallocated device buffer) or, fused, gsm_smem_collect_gathered writing each rank's ordered
records straight into the gathering rank's HBM through a peer mapping (NVLink / NVSwitch).
There is no data-path collective besides this one: every rank searches its own contiguous block of reads.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _capi as capi
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


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class NativeComm:
    """The library's NCCL communicator (gsm_comm_*), bootstrapped over an existing torch.distributed group: rank 0
    draws the 128-byte id, the group broadcasts it, every rank calls gsm_comm_init on its current CUDA device."""

    def __init__(self, group=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ident = np.zeros(128, np.uint8)
        if self.rank == 0:
            capi.check(capi.lib.gsm_comm_unique_id(ident.ctypes.data))
        box = [ident.tobytes()]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = np.frombuffer(box[0], np.uint8).copy()
        h = C.c_void_p()
        capi.check(capi.lib.gsm_comm_init(ident.ctypes.data, self.rank, self.world, C.byref(h)))
        self._h = h
        ver = C.c_int()
        capi.check(capi.lib.gsm_comm_info(self._h, None, None, C.byref(ver)))
        self.nccl_version = int(ver.value)

    def allgather_u64(self, send_dev, recv_dev, n=1):
        capi.check(capi.lib.gsm_comm_allgather_u64(self._h, C.c_void_p(send_dev), n, C.c_void_p(recv_dev), _stream()))

    def gather_records(self, send, n_send, counts, recv, dst=0):
        """send / recv: device uint8 tensors (or None); counts: host uint64 array, records per rank."""
        counts = np.ascontiguousarray(counts, np.uint64)
        capi.check(capi.lib.gsm_gather_records(self._h, C.c_void_p(send.data_ptr() if send is not None else 0), int(n_send),
                                               counts.ctypes.data, C.c_void_p(recv.data_ptr() if recv is not None else 0), int(dst), _stream()))

    def close(self):
        if self._h is not None:
            capi.lib.gsm_comm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_comms = {}


def native_comm(group=None):
    """One cached NativeComm per process group (collective on first use)."""
    key = id(group)
    if key not in _comms:
        _comms[key] = NativeComm(group)
    return _comms[key]


def gather_records(records, counts_per_read, dst=0, group=None, device=None, out=None):
    """Gather variable-length record arrays to `dst`.

    records: this rank's records (structured RECORD_DTYPE array, or a uint8 tensor of 16-byte records already on the
    device), read ids already global (read_id_base); counts_per_read: records per local read (int64 array / tensor).
    Returns (records, counts) concatenated in rank order on dst -- device tensors (uint8 view of 16-byte records, int64)
    when the backend is NCCL, numpy arrays with gloo -- and (None, None) elsewhere.  `out`: optional preallocated
    device uint8 tensor for the records on dst (NCCL).  Exact sizes travel: one all_gather of the two sizes, then
    gsm_gather_records / gsm-free torch point-to-point (gloo).  Nothing is padded, nothing is staged through the host.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nccl = dist.get_backend(group) == "nccl"
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    rec_u8 = torch.from_numpy(np.ascontiguousarray(records).view(np.uint8).copy()) if not torch.is_tensor(records) else records.view(torch.uint8)
    cnt = torch.from_numpy(np.ascontiguousarray(counts_per_read, dtype=np.int64)) if not torch.is_tensor(counts_per_read) else counts_per_read.to(torch.int64)
    rec_u8, cnt = rec_u8.to(device).contiguous(), cnt.to(device).contiguous()
    sizes = torch.tensor([rec_u8.numel() // 16, cnt.numel()], dtype=torch.int64, device=device)
    lst = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(lst, sizes, group=group)
    all_sizes = torch.stack(lst).cpu().numpy()
    n_rec, n_cnt = all_sizes[:, 0].astype(np.uint64), all_sizes[:, 1].astype(np.uint64)
    if nccl:
        comm = native_comm(group)
        recv_rec = recv_cnt = None
        if rank == dst:
            need = int(n_rec.sum()) * 16
            recv_rec = out[:need] if out is not None and out.numel() >= need else torch.empty(max(need, 16), dtype=torch.uint8, device=device)[:need]
            recv_cnt = torch.empty(max(int(n_cnt.sum()), 1), dtype=torch.int64, device=device)[: int(n_cnt.sum())]
        comm.gather_records(rec_u8, int(n_rec[rank]), n_rec, recv_rec, dst)
        # per-read counts: 8-byte items through the same exact-size exchange (two counts = one 16-byte "record")
        pad = cnt if cnt.numel() % 2 == 0 else torch.cat([cnt, cnt.new_zeros(1)])
        n_pair = ((n_cnt + np.uint64(1)) // np.uint64(2)).astype(np.uint64)
        recv_pairs = torch.empty(max(int(n_pair.sum()) * 2, 2), dtype=torch.int64, device=device) if rank == dst else None
        comm.gather_records(pad.view(torch.uint8), int(n_pair[rank]), n_pair, recv_pairs.view(torch.uint8) if recv_pairs is not None else None, dst)
        if rank != dst:
            return None, None
        a = 0
        parts = []
        for r in range(world):
            parts.append(recv_pairs[a:a + int(n_cnt[r])])
            a += int(n_pair[r]) * 2
        recv_cnt = torch.cat(parts) if parts else recv_cnt
        return recv_rec, recv_cnt
    # gloo (CPU tests): exact-size point-to-point
    if rank == dst:
        recs, cnts = [], []
        for r in range(world):
            if r == rank:
                recs.append(rec_u8)
                cnts.append(cnt)
                continue
            br = torch.empty(int(n_rec[r]) * 16, dtype=torch.uint8)
            bc = torch.empty(int(n_cnt[r]), dtype=torch.int64)
            if br.numel():
                dist.recv(br, src=r, group=group)
            if bc.numel():
                dist.recv(bc, src=r, group=group)
            recs.append(br)
            cnts.append(bc)
        return torch.cat(recs).numpy().view(RECORD_DTYPE), torch.cat(cnts).numpy()
    if rec_u8.numel():
        dist.send(rec_u8, dst=dst, group=group)
    if cnt.numel():
        dist.send(cnt, dst=dst, group=group)
    return None, None


class RecordGatherer:
    """The fused gather: every rank's ordered-write kernel stores its records directly into ONE buffer in the HBM of
    rank `dst` (peer mapping over NVLink / NVSwitch), batch after batch, with no host round trip, no separate copy and
    no collective per batch.

    Collective constructor: dst allocates `capacity` records, exports the allocation (CUDA IPC), the group carries the
    64-byte handle, the other ranks map it.  The buffer is cut into one REGION per rank (capacity / world records); rank r
    appends its batches to region r (gsm_smem_collect_gathered at region base + its own running count, kept on the device),
    so the ranks never wait for one another while they work.  fence(): ONE 8-byte-per-rank all-gather of the running counts
    -- dst's copy of it completes only after every rank's stream has passed its ordered writes -- and finish() returns, on
    dst, the per-rank views (rank order = global read order) and the counts; compact() joins them into one array.
    Without CUDA IPC (self.fused False) every batch falls back to gsm_gather_records (exact-size NCCL send / recv)."""

    def __init__(self, capacity, dst=0, group=None, device=None, comm=None):
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.comm = comm or native_comm(group)
        self.region = max(int(capacity) // self.world, 1)
        self.capacity = self.region * self.world
        self.base = torch.zeros(1, dtype=torch.int64, device=self.device)          # records this rank has written since reset()
        self.totals = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        self.n_batches = 0
        self._offset = 0
        handle = np.zeros(64, np.uint8)
        off = C.c_uint64(0)
        self.buffer = None
        self.out_ptr = 0
        ok, why = True, ""
        if self.rank == dst:
            self.buffer = torch.empty(self.capacity * 16, dtype=torch.uint8, device=self.device)
            try:
                capi.check(capi.lib.gsm_peer_export(C.c_void_p(self.buffer.data_ptr()), handle.ctypes.data, C.byref(off)))
            except Exception as e:          # e.g. an allocator whose memory cannot be exported
                ok, why = False, str(e)
        box = [(handle.tobytes(), int(off.value), ok, why)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
        ok, why = box[0][2], box[0][3]
        if ok and self.rank == dst:
            self.out_ptr = self.buffer.data_ptr()
        elif ok:
            handle = np.frombuffer(box[0][0], np.uint8).copy()
            p = C.c_void_p()
            try:
                capi.check(capi.lib.gsm_peer_open(handle.ctypes.data, box[0][1], C.byref(p)))
                self.out_ptr, self._offset = int(p.value), int(box[0][1])
            except Exception as e:
                ok, why = False, str(e)
        # every rank must take the same path: the peer mapping is used only if it worked everywhere
        flags = [None] * self.world
        dist.all_gather_object(flags, (bool(ok), why), group=group)
        self.fused = all(f[0] for f in flags)
        self.fallback_reason = next((f[1] for f in flags if not f[0]), "")
        if not self.fused and self.rank != dst and self.out_ptr:
            self.close()

    def reset(self):
        self.base.zero_()
        self.n_batches = 0

    def collect(self, engine, reads_c):
        """engine: the Engine whose workspace holds the batch just selected; reads_c: its gsm_dev_reads struct."""
        count = engine.counters.data_ptr() + 8                       # ws->counters[1]: records of this batch (device)
        if self.fused:
            region = self.out_ptr + self.rank * self.region * 16
            capi.check(capi.lib.gsm_smem_collect_gathered(C.byref(reads_c), C.byref(engine.ws), C.c_void_p(region), self.region,
                                                          None, 0, C.c_void_p(self.base.data_ptr()), _stream()))
        else:
            # no peer mapping (CUDA IPC unavailable): ordered write into the rank's own buffer, then exact-size NCCL send /
            # recv into the rank's region of dst's buffer -- needs the counts on the host: one all-gather + sync per batch
            capi.check(capi.lib.gsm_smem_collect(C.byref(reads_c), C.byref(engine.ws), C.c_void_p(engine.records.data_ptr()), engine.rec_cap, _stream()))
            mine = torch.stack([engine.counters[1], self.base[0]])                     # [records of this batch, records written so far]
            both = torch.empty(2 * self.world, dtype=torch.int64, device=self.device)
            self.comm.allgather_u64(mine.data_ptr(), both.data_ptr(), 2)
            both = both.cpu().numpy().reshape(self.world, 2)
            cnt, bases = both[:, 0].astype(np.uint64), both[:, 1]
            if int((bases + both[:, 0]).max()) > self.region:
                raise capi.GsmError(capi.E_CAPACITY, "RecordGatherer: a rank's region of the destination buffer is too small")
            for r in range(self.world):                    # one exact-size transfer per source rank, into that rank's region
                if self.rank == self.dst or self.rank == r:
                    one = np.zeros(self.world, np.uint64)
                    one[r] = cnt[r]
                    view = self.buffer[(r * self.region + int(bases[r])) * 16:] if self.rank == self.dst else None
                    self.comm.gather_records(engine.records, int(cnt[r]) if self.rank == r else 0, one, view, self.dst)
        capi.check(capi.lib.gsm_gather_advance(C.c_void_p(self.base.data_ptr()), C.c_void_p(count), 1, _stream()))
        self.n_batches += 1

    def fence(self):
        """Completion fence on the current stream (all ranks): dst's copy of this small all-gather of the running counts
        completes only after every rank's stream has reached it, i.e. after every rank's ordered writes into dst's buffer."""
        self.comm.allgather_u64(self.base.data_ptr(), self.totals.data_ptr())

    def finish(self):
        """Fence + synchronise.  On dst: (list of per-rank uint8 views of 16-byte records, in rank order; counts per rank)."""
        self.fence()
        torch.cuda.current_stream().synchronize()
        totals = self.totals.cpu().numpy()
        if int(totals.max()) > self.region:
            raise capi.GsmError(capi.E_CAPACITY, f"RecordGatherer: {int(totals.max())} records exceed a rank's region of {self.region}")
        if self.rank != self.dst:
            return None, totals
        return [self.buffer[r * self.region * 16:(r * self.region + int(totals[r])) * 16] for r in range(self.world)], totals

    def compact(self, parts):
        """One contiguous array of the gathered records in global read order (a device-to-device copy on dst)."""
        return torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device=self.device)

    def close(self):
        if self.rank != self.dst and self.out_ptr:
            capi.lib.gsm_peer_close(C.c_void_p(self.out_ptr), self._offset)
            self.out_ptr = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
