#!/usr/bin/env python3
"""Where does the end-to-end step spend its time?  PCIe rates of this box, pinned flags of the batch
buffers, and the pipelined path at several chunk counts (one JSON line each)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import genie_smem_b200 as g
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps({"copy": name, "gbytes_per_s": round(n / dt / 1e9, 1)}))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"copy": "h2d+d2h concurrent", "gbytes_per_s_each": round(n / dt / 1e9, 1)}))
    del h, d, h2, d2

    ref = bench.make_reference(100_000_000)
    packed = g.PackedIndex.from_host(g.HostIndex.build(bench._B[ref].tobytes()))
    index = g.DeviceIndex(packed, "cuda", with_sa=False, with_text=False)
    reads = bench.make_reads_host(ref, 10_000_000, bench.READ_LEN, seed=101)
    batch = g.ReadBatch.from_codes(reads, bench.READ_LEN, pin=True)
    print(json.dumps({"pinned": {k: bool(torch.from_numpy(v).is_pinned()) for k, v in
                                 (("packed", batch.packed_host), ("chunk_off", batch.chunk_off_host.view("int32")), ("len", batch.len_host.view("int32")))}}))
    batch.to("cuda")
    eng = g.Engine(index, batch.n, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
    for _ in range(2):
        eng.launch(g.METHOD_BWA, batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        eng.launch(g.METHOD_BWA, batch)
    torch.cuda.synchronize()
    print(json.dumps({"path": "device-resident launch", "ms": round((time.perf_counter() - t0) / 3 * 1e3, 2)}))
    for _ in range(2):
        eng.run(g.METHOD_BWA, batch)
    t0 = time.perf_counter()
    for _ in range(3):
        eng.run(g.METHOD_BWA, batch)
    print(json.dumps({"path": "Engine.run (serial H2D, kernels, D2H)", "ms": round((time.perf_counter() - t0) / 3 * 1e3, 2)}))
    del eng
    for nc in (2, 4, 8, 16):
        pipe = g.PipelinedEngine(index, batch.n, bench.READ_LEN, n_chunks=nc, mems_per_read=24, recs_per_read=8)
        for _ in range(2):
            pipe.run(g.METHOD_BWA, batch)
        t0 = time.perf_counter()
        for _ in range(3):
            pipe.run(g.METHOD_BWA, batch)
        print(json.dumps({"path": f"PipelinedEngine n_chunks={nc}", "ms": round((time.perf_counter() - t0) / 3 * 1e3, 2)}))
        del pipe


if __name__ == "__main__":
    main()
