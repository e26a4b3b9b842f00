#!/usr/bin/env python3
"""k_sweep time vs seed-table K on the bench's index shape (1 Gbp, 10 M reads)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import genie_smem_b200 as g  # noqa: E402

n_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
ks = [int(x) for x in sys.argv[3:]] or [0, 11, 12, 13, 14]
ref = bench.make_reference(n_ref, 1000)
index = g.DeviceIndex.build_on_device(ref, "cuda")
reads = bench.make_reads_host(ref, n_reads, bench.READ_LEN, seed=1001)
batch = g.ReadBatch.from_codes(reads, bench.READ_LEN).to("cuda")
eng = g.Engine(index, n_reads, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
for K in ks:
    if K:
        index.build_seed_table(K)
    else:
        index.drop_seed_table()
    for _ in range(2):
        eng.sweep(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.sweep(batch)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"seed_K": K, "table_MB": round((4 ** K) * 16 / 1e6) if K else 0, "sweep_ms": round(e0.elapsed_time(e1) / 3, 2),
                      "mems": eng.check_overflow()[0]}), flush=True)
