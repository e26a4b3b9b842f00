#!/usr/bin/env python3
"""Small fixed workload for ncu: the bench's step (BWA/LUT/RMI launch) on the bench's index shape
with fewer reads.  The packed index is cached under --cache (a directory on the box, e.g.
/dev/shm) so that the plain run and the ncu run of one gpurun call build it only once.

Usage: python tools/profile_step.py [--reads N] [--ref-bases B] [--method bwa|lut|rmi] [--cache DIR]"""
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
    ap.add_argument("--cache", default="")
    a = ap.parse_args()
    import torch
    import genie_smem_b200 as g
    t0 = time.time()
    ref = bench.make_reference(a.ref_bases)
    cdir = os.path.join(a.cache, f"gsm_index_{a.ref_bases}") if a.cache else ""
    if cdir and os.path.exists(os.path.join(cdir, "meta.json")):
        packed = g.PackedIndex.load(cdir, mmap=False)
    else:
        packed = g.PackedIndex.from_host(g.HostIndex.build(bench._B[ref].tobytes()))
        if cdir:
            packed.save(cdir)
    index = g.DeviceIndex(packed, "cuda")
    reads = bench.make_reads_host(ref, a.reads, bench.READ_LEN, seed=101)
    batch = g.ReadBatch.from_codes(reads, bench.READ_LEN).to("cuda")
    eng = g.Engine(index, a.reads, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
    kw = {}
    method = g.METHOD_BWA
    if a.method == "lut":
        method, kw = g.METHOD_LUT, {"K": bench.LUT_K, "lut": g.lut_build(index, bench.LUT_K)}
    elif a.method == "rmi":
        experts = bench.RMI_EXPERTS if a.ref_bases < 500_000_000 else (2048, 1048576)
        rmi = bench.train_rmi(packed.sa, ref, bench.RMI_K, experts, "cuda").build_probe_table(index)
        method, kw = g.METHOD_RMI, {"rmi": rmi}
    torch.cuda.synchronize()
    print(f"setup {time.time()-t0:.1f}s", file=sys.stderr)
    for _ in range(1 + a.steps):
        eng.launch(method, batch, **kw)
    torch.cuda.synchronize()
    print("mems, records:", eng.check_overflow())


if __name__ == "__main__":
    main()
