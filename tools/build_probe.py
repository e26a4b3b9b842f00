#!/usr/bin/env python3
"""Time the device index builder (gsm_index_build_device) at given reference sizes; optionally verify against
the host SA-IS builder.  python tools/build_probe.py 100000000 1000000000 [--verify-upto 20000000]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import genie_smem_b200 as g  # noqa: E402

sizes = [int(a) for a in sys.argv[1:] if not a.startswith("--")]
verify_upto = 20_000_000
if "--verify-upto" in sys.argv:
    verify_upto = int(sys.argv[sys.argv.index("--verify-upto") + 1])
for n in sizes:
    ref = np.random.Generator(np.random.PCG64(100)).integers(0, 4, n, dtype=np.uint8)
    torch.cuda.synchronize()
    t0 = time.time()
    idx = g.DeviceIndex.build_on_device(ref)
    torch.cuda.synchronize()
    wall = time.time() - t0
    out = {"n_bases": n, "wall_s": round(wall, 3), **{k: (round(v, 2) if isinstance(v, float) else v) for k, v in idx.build_stats.items()},
           "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}
    if n <= verify_upto:
        t0 = time.time()
        host = g.HostIndex.build(np.frombuffer(b"ACGT", np.uint8)[ref].tobytes())
        out["host_build_s"] = round(time.time() - t0, 2)
        fwd, rev, sa, _ = host.pack(with_sa=True, with_text=False)
        out["equal"] = bool(np.array_equal(idx.sa.cpu().numpy().view(np.uint32), sa) and np.array_equal(idx.fwd.cpu().numpy().view(np.uint32), fwd)
                            and np.array_equal(idx.rev.cpu().numpy().view(np.uint32), rev))
    print(json.dumps(out), flush=True)
    del idx
    torch.cuda.empty_cache()
