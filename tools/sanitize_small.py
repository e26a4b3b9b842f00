#!/usr/bin/env python3
"""Small all-kernel workload for compute-sanitizer (memcheck / racecheck): 1 Mbp reference, 3,000 reads of 151 bases (1 %
substitutions, uniform random, poly-A), the sweep and the three selections through Engine.run and PipelinedEngine.run, RMI
lookups from the bounds table, the seed table and the probe phases.  Prints record checksums (must not depend on the tool).

Usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import genie_smem_b200 as g
    L = 151
    rng = np.random.default_rng(77)
    ref = rng.integers(0, 4, 1_000_000, dtype=np.uint8)
    starts = rng.integers(0, len(ref) - L + 1, 3000)
    reads = ref[starts[:, None] + np.arange(L)[None, :]].copy()
    mut = rng.random(reads.shape) < 0.01
    reads[mut] = (reads[mut] + rng.integers(1, 4, int(mut.sum()), dtype=np.uint8)) & 3
    reads[:400] = rng.integers(0, 4, (400, L), dtype=np.uint8)
    reads[400:430, :70] = 0
    reads[430:440] = 0
    reads[440:520, 40:120] = np.tile(np.array([0, 1], np.uint8), 40)       # low complexity: long match lists
    idx = g.DeviceIndex.build_on_device(ref).build_seed_table(8)
    batch = g.ReadBatch.from_codes(reads, L)
    e = g.Engine(idx, len(reads), L, mems_per_read=256, recs_per_read=96)
    out = {}

    def sha(res):
        return hashlib.sha256(res.records.tobytes() + res.offsets.tobytes() + res.status.tobytes()).hexdigest()[:12]
    out["bwa"] = sha(e.run(g.METHOD_BWA, batch, min_len=1))
    out["bwa20"] = sha(e.run(g.METHOD_BWA, batch, min_len=20))
    out["lut"] = sha(e.run(g.METHOD_LUT, batch, K=8, lut=g.lut_build(idx, 8)))
    rmi = bench.train_rmi(idx, 9, (8, 256), idx.device)
    out["rmi_seed"] = sha(e.run(g.METHOD_RMI, batch, rmi=rmi))
    rmi.build_bounds_table(idx)
    out["rmi_bounds"] = sha(e.run(g.METHOD_RMI, batch, rmi=rmi))
    rmi.drop_bounds_table()
    idx.seed_table, idx.seed_K = None, 0
    idx._bind()
    out["rmi_probe"] = sha(e.run(g.METHOD_RMI, batch, rmi=rmi))
    out["bwa_noseed"] = sha(e.run(g.METHOD_BWA, batch, min_len=1))
    idx.build_seed_table(8)
    pipe = g.PipelinedEngine(idx, max_reads=len(reads), max_len=L, n_chunks=3, mems_per_read=256, recs_per_read=96)
    host = torch.from_numpy(np.frombuffer(b"ACGT", np.uint8)[reads].copy()).pin_memory()
    out["pipe_ascii"] = sha(pipe.run_ascii(g.METHOD_BWA, host, L, min_len=1))
    torch.cuda.synchronize()
    assert out["bwa"] == out["bwa_noseed"] == out["pipe_ascii"], out
    assert out["rmi_seed"] == out["rmi_bounds"] == out["rmi_probe"], out
    print(out)


if __name__ == "__main__":
    main()
