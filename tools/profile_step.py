#!/usr/bin/env python3
"""Small fixed workload for ncu: the bench's step (BWA/LUT/RMI launch) on the bench's index shape
(index built on the GPU, seed table as in bench.py) with fewer reads.

Usage: python tools/profile_step.py [--reads N] [--ref-bases B] [--method bwa|lut|rmi] [--seed-k K]"""
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
    ap.add_argument("--ref-bases", type=int, default=1_000_000_000)
    ap.add_argument("--method", default="bwa")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--seed-k", type=int, default=-1)
    ap.add_argument("--bounds", action="store_true", help="RMI: lookups from the dense k-mer bounds table")
    a = ap.parse_args()
    import torch
    import genie_smem_b200 as g
    t0 = time.time()
    seed = 1000 if a.ref_bases >= 500_000_000 else 100
    ref = bench.make_reference(a.ref_bases, seed)
    index = g.DeviceIndex.build_on_device(ref, "cuda")
    if a.seed_k != 0:
        index.build_seed_table(None if a.seed_k < 0 else a.seed_k)
    reads = bench.make_reads_host(ref, a.reads, bench.READ_LEN, seed=seed + 1)
    batch = g.ReadBatch.from_codes(reads, bench.READ_LEN).to("cuda")
    eng = g.Engine(index, a.reads, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
    kw = {}
    method = g.METHOD_BWA
    if a.method == "lut":
        method, kw = g.METHOD_LUT, {"K": bench.LUT_K, "lut": g.lut_build(index, bench.LUT_K)}
    elif a.method == "rmi":
        experts = bench.CONFIGS["c3" if a.ref_bases < 500_000_000 else "c4"]["experts"]
        method, kw = g.METHOD_RMI, {"rmi": bench.train_rmi(index, bench.RMI_K, experts, "cuda", probe_table=False)}
        if a.bounds:
            kw["rmi"].build_bounds_table(index)
    torch.cuda.synchronize()
    print(f"setup {time.time()-t0:.1f}s, seed table K={index.seed_K}", file=sys.stderr)
    for _ in range(1 + a.steps):
        eng.launch(method, batch, **kw)
    torch.cuda.synchronize()
    print("mems, records:", eng.check_overflow())


if __name__ == "__main__":
    main()
