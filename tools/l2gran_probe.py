#!/usr/bin/env python3
"""Effect of cudaLimitMaxL2FetchGranularity (32/64/128 B) on the random 64-byte gather ceiling and on k_sweep.
One process per setting (the limit is set before the context does any work)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    gran = int(sys.argv[2])
    import torch
    import genie_smem_b200 as g
    import bench
    torch.zeros(1, device="cuda")
    cur = g.l2_fetch_granularity(gran)
    out = {"requested": gran, "in_force": cur}
    for mb in (667, 4096):
        buf = torch.randint(0, 2**31 - 1, (mb * 1024 * 1024 // 4,), dtype=torch.int32, device="cuda")
        for dep in (0, 1):
            done, sec = g.gather_probe(buf, 400_000_000 if dep == 0 else 100_000_000, dep)
            out[f"gather_{mb}mb_{'dep' if dep else 'indep'}_gfetch_s"] = round(done / sec / 1e9, 2)
        del buf
    n_ref, n_reads = 1_000_000_000, 10_000_000
    ref = bench.make_reference(n_ref, 1000)
    index = g.DeviceIndex.build_on_device(ref, "cuda").build_seed_table()
    reads = bench.make_reads_host(ref, n_reads, bench.READ_LEN, seed=1001)
    batch = g.ReadBatch.from_codes(reads, bench.READ_LEN).to("cuda")
    eng = g.Engine(index, n_reads, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
    for _ in range(2):
        eng.sweep(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.sweep(batch)
    e1.record()
    torch.cuda.synchronize()
    out["sweep_ms_per_10M_reads"] = round(e0.elapsed_time(e1) / 3, 2)
    print(json.dumps(out), flush=True)
else:
    for gran in (0, 32, 64, 128):
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(gran)])
