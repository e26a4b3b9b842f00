#!/usr/bin/env python3
"""RMI-SMEM pre-filter on the device: same records with and without it, and what it buys.

A reference of --ref-bases bases with K chosen so that 4^K ~ rows (the code density of BASELINE.json configs[3]: K = 15 at
1 Gbp), --reads bench-style reads (151 bp, 1 % substitutions).  Prints one JSON line: hazard codes of the trained model, reads
left to the frame machine, select + scan + write time with the pre-filter and with the frame machine on every read (CUDA
events, after the same sweep), and whether the two record arrays are identical.

Usage: python tools/rmi_prefilter_check.py [--ref-bases B] [--reads N] [--K k]"""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bases", type=int, default=16_000_000)
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--K", type=int, default=12)
    ap.add_argument("--experts", type=int, nargs=2, default=(128, 16384))
    a = ap.parse_args(argv)
    import torch
    import genie_smem_b200 as g
    ref = bench.make_reference(a.ref_bases, 100)
    ref_dev = torch.from_numpy(ref).cuda()
    index = g.DeviceIndex.build_on_device(ref_dev, "cuda")
    index.build_seed_table(None)
    codes = torch.empty((a.reads, bench.READ_LEN), dtype=torch.uint8, device="cuda")
    bench.device_reads(ref_dev, a.reads, bench.READ_LEN, 101, codes)
    batch = g.ReadBatch.from_device_bases(codes, bench.READ_LEN)
    del codes, ref_dev
    eng = g.Engine(index, a.reads, bench.READ_LEN, mems_per_read=24, recs_per_read=8)
    rmi = bench.train_rmi(index, a.K, tuple(a.experts), "cuda", probe_table=False, bounds_table=True)
    out = {"ref_bases": a.ref_bases, "reads": a.reads, "K": a.K, "experts": list(a.experts), "hazard_codes": rmi.n_hazards,
           "codes": 4 ** a.K, "filter_active": rmi.hazard_slots is not None, "max_err_rows": getattr(rmi, "max_err", None)}

    def timed(fn, steps=3):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    def digest():
        torch.cuda.synchronize()
        _, n_rec = eng.check_overflow()
        h = hashlib.sha1()
        h.update(eng.records[: n_rec * 16].cpu().numpy().tobytes())
        h.update(eng.rec_off[: a.reads + 1].cpu().numpy().tobytes())
        h.update(eng.read_status[: a.reads].cpu().numpy().tobytes())
        return n_rec, h.hexdigest()

    eng.sweep(batch)
    out["ms_sweep"] = round(timed(lambda: eng.sweep(batch)), 3)
    for name, on in (("prefilter", True), ("frame_machine", False)):
        g.set_rmi_prefilter(on)
        out[f"ms_select_{name}"] = round(timed(lambda: eng.select(g.METHOD_RMI, batch, rmi=rmi)), 3)
        out[f"records_{name}"], out[f"sha_{name}"] = digest()
    g.set_rmi_prefilter(True)
    out["ms_select_bwa"] = round(timed(lambda: eng.select(g.METHOD_BWA, batch, min_len=1)), 3)
    out["identical"] = out["sha_prefilter"] == out["sha_frame_machine"]
    print(json.dumps(out), flush=True)
    return 0 if out["identical"] else 1


if __name__ == "__main__":
    sys.exit(main())
