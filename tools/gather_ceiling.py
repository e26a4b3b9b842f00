#!/usr/bin/env python3
"""Random-gather ceilings of this GPU with the kernels' access shapes (gsm_gather_probe2): G fetches/s and GB/s for
64 / 32 / 16-byte fetches, 1 / 4 / 8 fetches in flight per lane, over buffers from L2-resident to far beyond L2.
Run it under `ncu --metrics dram__bytes_read.sum,lts__t_sectors_op_read.sum` to see the DRAM bytes moved per fetch."""
import json
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import genie_smem_b200 as g

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 667, 4300]
out = []
for mb in sizes:
    buf = torch.randint(0, 2**31 - 1, (mb * 1024 * 1024 // 4,), dtype=torch.int32, device="cuda")
    for fb in (64, 66, 32, 16):
        for u in ((4,) if fb == 66 else (1, 4)):
            done, sec = g.gather_probe2(buf, 300_000_000, fb, u)
            out.append({"buffer_MB": mb, "fetch_bytes": fb, "in_flight": u, "gfetch_s": round(done / sec / 1e9, 2), "GBs": round(done * (64 if fb == 66 else fb) / sec / 1e9, 1), "note": {66: "64 B by a lane pair: one request"}.get(fb, "")})
            print(json.dumps(out[-1]), flush=True)
    del buf
