"""world_size-2 gloo test of the multi-GPU plumbing: reads shard by rank, records gather to rank 0
(the N>1 path of bench.py / north_star (4)), with the oracle standing in for the per-rank search."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, reads, text, sa, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from genie_smem_b200 import sharding
    from genie_smem_b200.engine import RECORD_DTYPE
    from oracle.c_oracle import COracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(len(reads), rank, world)
    mine = reads[lo:hi]
    out, counts = COracle(text, sa).smems(0, mine, min_len=1, threads=1)
    recs = np.zeros(int(counts.sum()), RECORD_DTYPE)
    k = 0
    for r in range(len(mine)):
        for o in out[r, :counts[r]]:
            recs[k] = (lo + r, o[0], o[1], o[2], o[3])
            k += 1
    g_recs, g_cnts = sharding.gather_records(recs, counts.astype(np.int64), dst=0)
    if rank == 0:
        ret["recs"] = g_recs.copy()
        ret["cnts"] = g_cnts.copy()
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    from tests import golden_util as gu
    from oracle.c_oracle import COracle
    from genie_smem_b200 import sharding
    g = gu.load_index("medium_data")
    rng = np.random.default_rng(2)
    reads = []
    for _ in range(41):                      # odd count: ragged shards
        L = int(rng.integers(20, 120))
        p = int(rng.integers(0, len(g["text"]) - L))
        reads.append(g["text"][p:p + L])
    assert [sharding.shard_range(41, r, 2) for r in range(2)] == [(0, 20), (20, 41)]
    assert [sharding.shard_range(10, r, 8) for r in range(8)][-1] == (8, 10)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, reads, g["text"], g["suffix_array"], ret), nprocs=2, join=True)
    out, counts = COracle(g["text"], g["suffix_array"]).smems(0, reads, min_len=1, threads=1)
    assert np.array_equal(ret["cnts"], counts.astype(np.int64))
    recs = ret["recs"]
    k = 0
    for r in range(len(reads)):
        for o in out[r, :counts[r]]:
            assert tuple(int(x) for x in recs[k]) == (r, int(o[0]), int(o[1]), int(o[2]), int(o[3]))
            k += 1
    assert k == len(recs)
