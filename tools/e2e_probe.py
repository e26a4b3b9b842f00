#!/usr/bin/env python3
"""End-to-end path (raw ASCII reads in pinned host memory -> host records) vs chunk count, next to the device-resident step.
python tools/e2e_probe.py [ref_bases] [reads] [chunk counts...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import genie_smem_b200 as g  # noqa: E402

n_ref = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
chunks = [int(x) for x in sys.argv[3:]] or [8, 16, 32]
dev = torch.device("cuda")
ref = bench.make_reference(n_ref, 1000)
ref_dev = torch.from_numpy(ref).to(dev)
index = g.DeviceIndex.build_on_device(ref_dev, dev).build_seed_table()
codes = torch.empty((n_reads, bench.READ_LEN), dtype=torch.uint8, device=dev)
bench.device_reads(ref_dev, n_reads, bench.READ_LEN, 1001, codes)
batch = g.ReadBatch.from_device_bases(codes, bench.READ_LEN)
lut4 = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
ascii_host = torch.empty((n_reads, bench.READ_LEN), dtype=torch.uint8, pin_memory=True)
for a in range(0, n_reads, 5_000_000):
    ascii_host[a:a + 5_000_000].copy_(lut4[codes[a:a + 5_000_000].long()])
del codes, ref_dev
eng = g.Engine(index, n_reads, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
for _ in range(2):
    eng.launch(g.METHOD_BWA, batch, min_len=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.launch(g.METHOD_BWA, batch, min_len=1)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"path": "device-resident launch", "ms": round(e0.elapsed_time(e1) / 3, 2)}), flush=True)
del eng
torch.cuda.empty_cache()
for nc in chunks:
    pipe = g.PipelinedEngine(index, n_reads, bench.READ_LEN, n_chunks=nc, mems_per_read=24, recs_per_read=8)
    for _ in range(2):
        pipe.run_ascii(g.METHOD_BWA, ascii_host, bench.READ_LEN, min_len=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        res = pipe.run_ascii(g.METHOD_BWA, ascii_host, bench.READ_LEN, min_len=1)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print(json.dumps({"path": f"PipelinedEngine.run_ascii n_chunks={nc}", "ms": round(ms, 2), "Mreads_per_s": round(n_reads / ms / 1e3, 1),
                      "launch_chunks": len(pipe._chunk_bounds(n_reads)) - 1, "records": int(len(res.records))}), flush=True)
    del pipe
    torch.cuda.empty_cache()
