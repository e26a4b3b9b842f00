#!/usr/bin/env python3
"""A/B of the selection kernels (GSM_SELECT_TEAMS) on the bench's index shape: select-only ms per method and mode.
One child process per mode (the switch is read once per process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import bench
    import genie_smem_b200 as g
    n_ref, n_reads = int(sys.argv[2]), int(sys.argv[3])
    ref = bench.make_reference(n_ref, 1000)
    index = g.DeviceIndex.build_on_device(ref, "cuda").build_seed_table()
    reads = bench.make_reads_host(ref, n_reads, bench.READ_LEN, seed=1001)
    batch = g.ReadBatch.from_codes(reads, bench.READ_LEN).to("cuda")
    eng = g.Engine(index, n_reads, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
    lut = g.lut_build(index, bench.LUT_K)
    rmi = bench.train_rmi(index, bench.RMI_K, bench.CONFIGS["c4"]["experts"] if n_ref >= 500_000_000 else bench.CONFIGS["c3"]["experts"], "cuda")
    if os.environ.get("GSM_RMI_PERSIST") == "1":
        rmi.persist_in_l2()
    eng.sweep(batch)
    out = {"GSM_SELECT_TEAMS": os.environ.get("GSM_SELECT_TEAMS", "0"), "GSM_RMI_PERSIST": os.environ.get("GSM_RMI_PERSIST", "0")}
    sums = {}
    for name, method, kw in (("bwa", g.METHOD_BWA, {"min_len": 1}), ("lut", g.METHOD_LUT, {"K": bench.LUT_K, "lut": lut}), ("rmi", g.METHOD_RMI, {"rmi": rmi})):
        for _ in range(2):
            eng.select(method, batch, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.select(method, batch, **kw)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_select_ms"] = round(e0.elapsed_time(e1) / 3, 2)
        n_mems, n_rec = eng.check_overflow()
        sums[name] = (n_rec, int(eng.records[: n_rec * 16].view(torch.int32).to(torch.int64).sum().item()))
    out["checksums"] = sums
    print(json.dumps(out), flush=True)
else:
    n_ref = sys.argv[1] if len(sys.argv) > 1 else "1000000000"
    n_reads = sys.argv[2] if len(sys.argv) > 2 else "10000000"
    for mode in ("0", "3"):
        env = dict(os.environ, GSM_SELECT_TEAMS=mode)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", n_ref, n_reads], env=env)
