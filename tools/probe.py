#!/usr/bin/env python3
"""Random aligned 64-byte gather ceilings on this GPU: the denominator next to the
streaming-copy peak.  Prints one JSON line per (buffer size, dependent?) point."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import genie_smem_b200 as g
    props = torch.cuda.get_device_properties(0)
    print(json.dumps({"gpu": props.name, "sms": props.multi_processor_count, "l2_bytes": props.L2_cache_size}))
    for mb in (32, 64, 512, 4096):
        buf = torch.randint(0, 2**31 - 1, (mb * 1024 * 1024 // 4,), dtype=torch.int32, device="cuda")
        for dep in (0, 1):
            n = 400_000_000 if dep == 0 else 100_000_000
            done, sec = g.gather_probe(buf, n, dep)
            print(json.dumps({"buffer_mb": mb, "dependent_chains": bool(dep), "fetches": done, "seconds": round(sec, 5),
                              "gfetch_per_s": round(done / sec / 1e9, 2), "gbytes_per_s": round(done * 64 / sec / 1e9, 1)}))
        del buf


if __name__ == "__main__":
    main()
