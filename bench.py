#!/usr/bin/env python3
"""bench.py -- SMEM reads/s on B200 (BASELINE.json metric) for the B200-native engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--ref-bases B] [--impl reference]

Workload at N=1 = BASELINE.json configs[2]: synthetic 100 Mbp random ACGT reference
(numpy PCG64(100)), 10 M reads of 151 bp (exact substrings with i.i.d. 1 % substitutions,
SURVEY 8d), all three SMEM methods on one B200.  Under torchrun (N>1) every rank holds a replica
of the index and its own 10 M-read shard (weak scaling); the only collective is the final gather
of per-rank record counts (NCCL), as north_star prescribes.

A "step" is one pass of the hot path over one read batch.  `value` = BWA-SMEM reads/s with the
packed reads already resident in HBM (kernels only: sweep + select + scan + gather); `e2e` = the
same through the public API with HOST buffers (pinned H2D of the packed reads, D2H of records and
offsets inside the timed region).  LUT- and RMI-SMEM throughputs are reported under "methods".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 151
SUB_RATE = 0.01
LUT_K = 12
RMI_K = 15
RMI_EXPERTS = (512, 131072)      # ~1 leaf per 760 keys, the reference's [10,100] ratio on its 100 kb text
_B = np.frombuffer(b"ACGT", dtype=np.uint8)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_reference(n_bases, seed=100):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 4, n_bases, dtype=np.uint8)


def make_reads_host(ref_codes, n_reads, L, seed, sub_rate=SUB_RATE, chunk=1_000_000):
    """(n_reads, L) uint8 codes: exact substrings with i.i.d. substitutions (seed = config seed + 1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n_reads, L), np.uint8)
    ar = np.arange(L, dtype=np.int64)[None, :]
    for a in range(0, n_reads, chunk):
        b = min(n_reads, a + chunk)
        starts = rng.integers(0, len(ref_codes) - L + 1, b - a)
        r = ref_codes[starts[:, None] + ar]
        mut = rng.random(r.shape) < sub_rate
        r[mut] = (r[mut] + rng.integers(1, 4, int(mut.sum()), dtype=np.uint8)) & 3
        out[a:b] = r
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(text_bytes, sa_1based, reads_codes, method, threads, budget_s, K=0, rmi=None):
    """Time the oracle port (oracle/smem_oracle.c, the reference algorithm restated in C) on the host
    cores over a bounded sample of the same reads.  Returns (reads/s, n_sample, threads)."""
    from oracle.c_oracle import COracle
    o = COracle(text_bytes, sa_1based)
    threads = threads or o.max_threads
    L = reads_codes.shape[1]

    def run(n):
        sub = reads_codes[:n]
        joined = _B[sub.reshape(-1)].tobytes()
        lens = np.full(n, L, np.uint32)
        t0 = time.perf_counter()
        o.smems(method, None, min_len=1, K=K, rmi=rmi, threads=threads, joined=joined, lens=lens)
        return time.perf_counter() - t0

    n = min(64 * threads, len(reads_codes))
    if method == 1:
        o.lib.orc_build_lut(o.h, K)          # table build is index construction, not search
    dt = run(n)
    rate = n / dt
    n2 = int(max(n, min(len(reads_codes), rate * budget_s)))
    if n2 > 2 * n:
        dt = run(n2)
        n = n2
    return n / dt, n, threads


def rmi_dict(params):
    return {"K": params.K, "level_sizes": [int(x) for x in params.level_sizes], "coef": params.coef_host, "intercept": params.intercept_host}


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--ref-bases", type=int, default=100_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget per method")
    ap.add_argument("--skip-rmi", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="chunks of the pipelined end-to-end path")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    global RMI_EXPERTS
    if args.ref_bases >= 500_000_000:           # keep ~1 leaf per 760-950 keys as the reference grows
        RMI_EXPERTS = (2048, 1048576)
    cfg_name = "BASELINE.json configs[2]" if args.ref_bases == 100_000_000 else \
        ("BASELINE.json configs[3] shape, full suffix array in HBM" if args.ref_bases == 1_000_000_000 else "custom size")
    workload = {"workload": f"synthetic {args.ref_bases/1e6:g} Mbp random ACGT reference (PCG64 seed 100), "
                            f"{args.reads/1e6:g} M reads x {READ_LEN} bp per GPU, exact substrings + {SUB_RATE:.0%} substitutions "
                            f"({cfg_name})",
                "ref_bases": args.ref_bases, "reads_per_gpu": args.reads, "read_len": READ_LEN, "sub_rate": SUB_RATE,
                "lut_K": LUT_K, "rmi_K": RMI_K, "rmi_experts": list(RMI_EXPERTS), "min_len": 1,
                "parallelism": f"reads sharded x{world}, index replicated", "l2_policy": f"packed read batch ({args.reads * 48 / 1e6:.0f} MB) and outputs exceed "
                f"L2; the rank buckets are {2 * (args.ref_bases // 192 + 1) * 64 / 1e6:.0f} MB (126 MB L2: resident at 100 Mbp, "
                "HBM-resident at 1 Gbp; see DESIGN.md)"}

    if args.impl == "reference":
        if rank != 0:
            return
        reference_arm(args, workload)
        return

    import torch
    import torch.distributed as dist
    import genie_smem_b200 as g
    from genie_smem_b200 import engine as eng

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    t_setup = time.time()
    ref = make_reference(args.ref_bases)
    text = _B[ref].tobytes()
    host = g.HostIndex.build(text)
    log(f"[rank {rank}] index built in {time.time()-t_setup:.1f}s")
    index = g.DeviceIndex(host, dev)
    reads_codes = make_reads_host(ref, args.reads, READ_LEN, seed=101 + rank)
    batch = g.ReadBatch.from_codes(reads_codes, READ_LEN, read_id_base=0, pin=True)
    batch.to(dev)
    engine = g.Engine(index, args.reads, READ_LEN, mems_per_read=24, recs_per_read=8)
    lut = g.lut_build(index, LUT_K)
    # RMI: trained on the host from this index (any (coef, intercept) set is valid input; the
    # oracle is given the same parameters)
    rmi = None
    if not args.skip_rmi:
        t0 = time.time()
        rmi = train_rmi(host, ref, RMI_K, RMI_EXPERTS, dev)
        rmi.build_probe_table(index)      # 16-byte {SA, 32-mer} probe records: one fetch per last-mile probe
        log(f"[rank {rank}] RMI trained in {time.time()-t0:.1f}s")
    torch.cuda.synchronize()
    log(f"[rank {rank}] setup {time.time()-t_setup:.1f}s, index {index.bytes()/1e6:.0f} MB on device")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- headline: BWA-SMEM, device-resident inputs
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_region0 = time.time()
    launches0 = engine.kernel_launches
    ms_bwa = timed(lambda: engine.launch(g.METHOD_BWA, batch, min_len=1), args.steps, args.warmup)
    gpu_launches = (engine.kernel_launches - launches0) * args.steps // (args.steps + args.warmup)
    t_region1 = time.time()
    clocks = sampler.stop(t_region0, t_region1)
    n_mems, n_rec = engine.check_overflow()

    # ---- per-kernel split of the step + algorithmic bytes (roofline of the dominant kernel, k_sweep)
    ms_sweep = timed(lambda: engine.sweep(batch), args.steps, 1)
    engine.sweep(batch)
    ms_sel_bwa = timed(lambda: engine.select(g.METHOD_BWA, batch, min_len=1), args.steps, 1)
    recs = engine.records[: n_rec * 16].view(torch.int32).view(-1, 4)
    qs = (recs[:, 1] & 0xFFFF).to(torch.int64)
    qe = ((recs[:, 1] >> 16) & 0xFFFF).to(torch.int64)
    steps_alg = int(((qe - qs) + (qs > 0).to(torch.int64)).sum().item())
    alg_bytes = 128 * steps_alg + args.reads * ((READ_LEN + 3) // 4) + 16 * n_rec
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_sweep * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_sweep_dram_bytes.json")
    if os.path.exists(tpath) and args.ref_bases == 100_000_000:      # the ncu capture was taken on this config
        try:
            tj = json.load(open(tpath))
            traffic = tj["dram_bytes_per_read"] * args.reads
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_sweep", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_read": round(alg_bytes / args.reads, 1), "fm_steps_per_read_min": round(steps_alg / args.reads, 2),
                "records_per_read": round(n_rec / args.reads, 3), "ms_sweep": round(ms_sweep, 3), "ms_select_scan_gather": round(ms_sel_bwa, 3),
                "note": (f"rank buckets = {2 * (args.ref_bases // 192 + 1) * 64 / 1e6:.0f} MB; below ~100 MB they stay in the 126 MB L2 and the "
                         "HBM fraction is a statement about necessary bytes, not a DRAM utilisation (see traffic)")}

    # ---- the other two methods (device-resident)
    methods = {"bwa": {"reads_per_s": world * args.reads / (ms_bwa * 1e-3), "ms_per_step": ms_bwa}}
    ms_lut = timed(lambda: engine.launch(g.METHOD_LUT, batch, K=LUT_K, lut=lut), max(2, args.steps // 2), 1)
    methods["lut"] = {"reads_per_s": world * args.reads / (ms_lut * 1e-3), "ms_per_step": ms_lut, "K": LUT_K}
    engine.check_overflow()
    if rmi is not None:
        ms_rmi = timed(lambda: engine.launch(g.METHOD_RMI, batch, rmi=rmi), 2, 1)
        methods["rmi"] = {"reads_per_s": world * args.reads / (ms_rmi * 1e-3), "ms_per_step": ms_rmi, "K": RMI_K,
                          "experts": list(RMI_EXPERTS)}
        st = engine.read_status[: args.reads]
        methods["rmi"]["reads_where_reference_raises"] = int((st == g.READ_REF_RAISES).sum().item())
        engine.check_overflow()

    # ---- end to end through the public API: pinned host reads in, host records out
    pipe = g.PipelinedEngine(index, args.reads, READ_LEN, n_chunks=args.e2e_chunks, mems_per_read=24, recs_per_read=8)

    def e2e_step():
        return pipe.run(g.METHOD_BWA, batch, min_len=1)      # pinned host reads in, host records out

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": world * args.reads * args.steps / dt, "unit": "reads/s", "h2d_bytes_per_step": batch.h2d_bytes(),
           "d2h_bytes_per_step": pipe.last_d2h_bytes, "method": "bwa", "records_last_step": int(len(res.records)),
           "api": f"PipelinedEngine.run ({args.e2e_chunks} chunks, 2 streams: H2D / kernels / D2H overlapped)"}

    # ---- multi-GPU: the one collective of the path -- gather of per-rank record counts to rank 0
    total_records = n_rec
    gather = None
    if world > 1:
        from genie_smem_b200 import sharding
        # the final gather of per-rank SMEM records to rank 0 over NCCL (device buffers, NVLink)
        rec_dev = engine.records[: n_rec * 16]
        cnt_dev = engine.rec_cnt[: args.reads].to(torch.int64)
        barrier()
        t0 = time.perf_counter()
        g_recs, g_cnts = sharding.gather_records(rec_dev, cnt_dev, dst=0, device=dev)
        torch.cuda.synchronize()
        dtg = time.perf_counter() - t0
        if rank == 0:
            total_records = int(len(g_recs))
            gather = {"ms": round(dtg * 1e3, 2), "bytes_received": int(len(g_recs)) * 16, "backend": "nccl",
                      "note": "includes the device-to-host read of the gathered array on rank 0; outside the timed steps"}

    # ---- CPU baseline on the host cores, rank 0 at N=1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        sa1, _ = host.export()
        v, n_s, thr = cpu_arm(text, sa1, reads_codes, 0, 0, args.cpu_seconds)
        cpu_baseline = {"value": round(v, 1), "unit": "reads/s", "cores": thr, "kind": "port",
                        "sample": f"first {n_s} reads of the same batch, BWA-SMEM, oracle/smem_oracle.c (reference algorithm, "
                                  f"O(L^2) restarts included) on {thr} pthreads"}
        v2, n2, _ = cpu_arm(text, sa1, reads_codes, 1, 0, args.cpu_seconds / 2, K=LUT_K)
        cpu_baseline["lut_reads_per_s"] = round(v2, 1)
        if rmi is not None:
            v3, n3, _ = cpu_arm(text, sa1, reads_codes, 2, 0, args.cpu_seconds / 2, rmi=rmi_dict(rmi))
            cpu_baseline["rmi_reads_per_s"] = round(v3, 1)

    if rank == 0:
        out = {"metric": "SMEM reads/sec (BWA-SMEM; LUT/RMI under methods)", "value": world * args.reads / (ms_bwa * 1e-3), "unit": "reads/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_bwa, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload,
               "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(gpu_launches),
               "clocks": clocks, "methods": methods, "records_total": total_records, "maximal_matches_per_read": round(n_mems / args.reads, 3),
               "record_gather": gather}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def train_rmi(host, ref_codes, K, experts, dev):
    """Host-side RMI training on the k-mer -> row keys of this index (reference RMI_LUT.py:36-50);
    every `stride`-th key to keep setup short -- model quality only changes the last-mile length."""
    import genie_smem_b200 as g
    sa1 = host.export()[0] if hasattr(host, "export") else np.asarray(host)     # a HostIndex, or the 1-based suffix array itself
    n_bases = len(ref_codes)
    stride = max(1, len(sa1) // 4_000_000)
    rows = np.arange(0, len(sa1), stride, dtype=np.int64)
    start = sa1[rows].astype(np.int64) - 1
    ok = start + K <= n_bases
    rows, start = rows[ok], start[ok]
    key = np.zeros(len(rows), np.int64)
    for j in range(K):
        key = (key << 2) | ref_codes[start + j]
    m = g.RMI(list(experts)).fit(key, rows)
    return g.RmiParams(K, m.level_sizes, m.coef, m.intercept, dev)


def reference_arm(args, workload):
    """--impl reference: the reference's CPU path (oracle port in C, all host threads) on bounded
    samples of the same workload.  No GPU, none of the engine's kernels."""
    import genie_smem_b200 as g        # host-side index builder only (SA-IS); search is the oracle's
    ref = make_reference(args.ref_bases)
    text = _B[ref].tobytes()
    host = g.HostIndex.build(text, reverse=False)
    sa1, _ = host.export()
    reads_codes = make_reads_host(ref, min(args.reads, 200_000), READ_LEN, seed=101)
    from oracle.c_oracle import COracle
    o = COracle(text, sa1)
    thr = o.max_threads
    L = READ_LEN
    # size one step to ~10 s
    n0 = 32 * thr
    joined = _B[reads_codes[:n0].reshape(-1)].tobytes()
    t0 = time.perf_counter()
    o.smems(0, None, min_len=1, threads=thr, joined=joined, lens=np.full(n0, L, np.uint32))
    rate = n0 / (time.perf_counter() - t0)
    n = int(min(len(reads_codes), max(n0, rate * 8.0)))
    joined = _B[reads_codes[:n].reshape(-1)].tobytes()
    lens = np.full(n, L, np.uint32)
    for _ in range(min(args.warmup, 1)):
        o.smems(0, None, min_len=1, threads=thr, joined=joined, lens=lens)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.smems(0, None, min_len=1, threads=thr, joined=joined, lens=lens)
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt
    sample = f"{n} reads per step (first reads of the batch), BWA-SMEM, oracle/smem_oracle.c on {thr} pthreads"
    out = {"impl": "reference", "metric": "SMEM reads/sec (BWA-SMEM; LUT/RMI under methods)", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload,
           "cpu_baseline": {"value": v, "unit": "reads/s", "cores": thr, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
