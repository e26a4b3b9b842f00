#!/usr/bin/env python3
"""Small fixed workload for ncu: the bench's step (BWA/LUT/RMI launch) on the bench's index shape
with fewer reads.  Usage: python tools/profile_step.py [--reads N] [--ref-bases B] [--method bwa|lut|rmi]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--ref-bases", type=int, default=100_000_000)
    ap.add_argument("--method", default="bwa")
    ap.add_argument("--steps", type=int, default=2)
    a = ap.parse_args()
    import torch
    import genie_smem_b200 as g
    t0 = time.time()
    ref = bench.make_reference(a.ref_bases)
    host = g.HostIndex.build(bench._B[ref].tobytes())
    index = g.DeviceIndex(host, "cuda")
    reads = bench.make_reads_host(ref, a.reads, bench.READ_LEN, seed=101)
    batch = g.ReadBatch.from_codes(reads, bench.READ_LEN).to("cuda")
    eng = g.Engine(index, a.reads, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
    kw = {}
    method = g.METHOD_BWA
    if a.method == "lut":
        method, kw = g.METHOD_LUT, {"K": bench.LUT_K, "lut": g.lut_build(index, bench.LUT_K)}
    elif a.method == "rmi":
        method, kw = g.METHOD_RMI, {"rmi": bench.train_rmi(host, ref, bench.RMI_K, bench.RMI_EXPERTS, "cuda")}
    torch.cuda.synchronize()
    print(f"setup {time.time()-t0:.1f}s", file=sys.stderr)
    for _ in range(1 + a.steps):
        eng.launch(method, batch, **kw)
    torch.cuda.synchronize()
    print("mems, records:", eng.check_overflow())


if __name__ == "__main__":
    main()
