#!/usr/bin/env python3
"""bench.py -- SMEM reads/s on B200 (BASELINE.json metric) for the B200-native engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c4|c3] [--reads R] [--ref-bases B] [--impl reference]

Workload at N=1 = BASELINE.json configs[3], the configuration the metric ("reads/s at 1/2/4/8 B200 + achieved HBM GB/s")
is quoted on: synthetic 1 Gbp random ACGT reference (numpy PCG64(1000)), 50 M reads of 151 bp (exact substrings with
i.i.d. 1 % substitutions), all three SMEM methods.  The index (suffix array, both BWT bucket arrays) is built
ON THE GPU in well under a second (gsm_index_build_device), so the 1 Gbp configuration fits the default run; `--config c3`
selects configs[2] (100 Mbp, 10 M reads; rank buckets L2-resident).  Under torchrun (N>1) every rank holds a replica of the
index and its own read shard (weak scaling); the only collective is the final gather of per-rank records (NCCL).

A "step" is one pass of the hot path over one read batch.  `value` = BWA-SMEM reads/s with the packed reads already
resident in HBM (kernels only: sweep + select + scan + gather); `e2e` = the same through the public API starting from RAW
read bytes in pinned host memory (1 byte/base H2D, 2-bit packing on the GPU, records + offsets D2H, all inside the timed
region).  LUT- and RMI-SMEM throughputs are reported under "methods".
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
_B = np.frombuffer(b"ACGT", dtype=np.uint8)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: keep a private copy of it and point fd 1 at stderr, so that nothing a library
# prints (NCCL's version banner under NCCL_DEBUG, torch warnings) can land next to the result
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_result(obj):
    _claim_stdout()
    print(json.dumps(obj), file=_RESULT_OUT, flush=True)


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
CONFIGS = {"c4": dict(ref_bases=1_000_000_000, reads=50_000_000, seed=1000, experts=(2048, 1048576), name="BASELINE.json configs[3]"),
           "c3": dict(ref_bases=100_000_000, reads=10_000_000, seed=100, experts=(512, 131072), name="BASELINE.json configs[2]")}
HOST_READS = 600_000      # the first reads of every batch come from numpy so the CPU arms can regenerate exactly them


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (default: the config's)")
    ap.add_argument("--ref-bases", type=int, default=0, help="reference size (default: the config's)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline budget per method")
    ap.add_argument("--skip-rmi", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="chunks of the pipelined end-to-end path")
    ap.add_argument("--seed-k", type=int, default=-1, help="K of the sweep kernel's seed table (-1 = auto, 0 = none)")
    ap.add_argument("--sub-rate", type=float, default=SUB_RATE, help="substitution rate of the synthetic reads (SURVEY 8d extremes: 0)")
    ap.add_argument("--random-reads", action="store_true", help="uniform random reads instead of reference substrings (SURVEY 8d extreme)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.ref_bases:
        cfg["ref_bases"] = args.ref_bases
        cfg["name"] = "custom size"
        if args.ref_bases < 500_000_000:
            cfg["experts"] = CONFIGS["c3"]["experts"]
    if args.reads:
        cfg["reads"] = args.reads
        if cfg["name"] != "custom size":
            cfg["name"] += f" shape, {args.reads/1e6:g} M reads"
    args.ref_bases, args.reads, args.seed, args.experts, args.cfg_name = cfg["ref_bases"], cfg["reads"], cfg["seed"], cfg["experts"], cfg["name"]
    return args


def workload_dict(args, world):
    bucket_mb = 2 * (args.ref_bases // 192 + 1) * 64 / 1e6
    return {"workload": f"synthetic {args.ref_bases/1e6:g} Mbp random ACGT reference (PCG64 seed {args.seed}), "
                        f"{args.reads/1e6:g} M reads x {READ_LEN} bp per GPU, "
                        + ("uniform random reads" if args.random_reads else f"exact substrings + {args.sub_rate:.0%} substitutions") + f" ({args.cfg_name})",
            "ref_bases": args.ref_bases, "reads_per_gpu": args.reads, "read_len": READ_LEN,
            "sub_rate": None if args.random_reads else args.sub_rate,
            "lut_K": LUT_K, "rmi_K": RMI_K, "rmi_experts": list(args.experts), "min_len": 1,
            "parallelism": f"reads sharded x{world}, index replicated",
            "l2_policy": f"inputs larger than L2: packed read batch {args.reads * 48 / 1e6:.0f} MB, rank buckets {bucket_mb:.0f} MB, outputs "
                         f"> 1 GB vs 126 MB L2 (at --config c3 the 67 MB of buckets are L2-resident by design; see DESIGN.md)"}


def host_reads(ref_codes, n, seed, sub_rate=SUB_RATE, random_reads=False):
    if random_reads:
        return np.random.Generator(np.random.PCG64(seed)).integers(0, 4, (n, READ_LEN), dtype=np.uint8)
    return make_reads_host(ref_codes, n, READ_LEN, seed=seed, sub_rate=sub_rate)


def device_reads(ref_dev, n, L, seed, out, sub_rate=SUB_RATE, chunk=2_000_000):
    """Fill out[(n, L) uint8, device] with synthetic reads: exact substrings + i.i.d. substitutions (torch RNG on the
    device; workload generation only)."""
    import torch
    gen = torch.Generator(device=ref_dev.device)
    gen.manual_seed(seed)
    ar = torch.arange(L, device=ref_dev.device, dtype=torch.int64)[None, :]
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        starts = torch.randint(0, ref_dev.numel() - L + 1, (b - a, 1), generator=gen, device=ref_dev.device, dtype=torch.int64)
        r = ref_dev[(starts + ar).view(-1)].view(b - a, L)
        mut = torch.rand(r.shape, generator=gen, device=ref_dev.device) < sub_rate
        add = torch.randint(1, 4, r.shape, generator=gen, device=ref_dev.device, dtype=torch.uint8)
        out[a:b] = torch.where(mut, (r + add) & 3, r)
    return out


def main():
    args = parse_args()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = workload_dict(args, world)

    if args.impl == "reference":
        if rank != 0:
            return
        reference_arm(args, workload)
        return

    import torch
    import torch.distributed as dist
    import genie_smem_b200 as g

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        from genie_smem_b200 import sharding
        cpus = sharding.bind_to_gpu_numa(local_rank)                # pinned host buffers land on the GPU's NUMA node
        log(f"[rank {rank}] bound to {len(cpus) if cpus else 'no'} CPUs next to GPU {local_rank}")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    # ---- setup: reference, index (built on the GPU), reads, LUT, RMI
    t_setup = time.time()
    ref = make_reference(args.ref_bases, args.seed)
    ref_dev = torch.from_numpy(ref).to(dev)
    t0 = time.time()
    index = g.DeviceIndex.build_on_device(ref_dev, dev)
    log(f"[rank {rank}] reference generated in {t0-t_setup:.1f}s; index built on the GPU in {index.build_stats['build_ms']:.0f} ms "
        f"(workspace {index.build_stats['workspace_bytes']/1e9:.1f} GB, {index.build_stats['doubling_rounds']} doubling rounds)")
    if args.seed_k != 0:
        index.build_seed_table(None if args.seed_k < 0 else args.seed_k)
    n_host = min(HOST_READS, args.reads)
    reads_head = host_reads(ref, n_host, args.seed + 1 + rank, args.sub_rate, args.random_reads)   # numpy: the CPU arms regenerate exactly these
    codes_dev = torch.empty((args.reads, READ_LEN), dtype=torch.uint8, device=dev)
    codes_dev[:n_host] = torch.from_numpy(reads_head).to(dev)
    if args.reads > n_host:
        if args.random_reads:
            gen = torch.Generator(device=dev)
            gen.manual_seed(args.seed + 1 + rank)
            codes_dev[n_host:] = torch.randint(0, 4, (args.reads - n_host, READ_LEN), generator=gen, device=dev, dtype=torch.uint8)
        else:
            device_reads(ref_dev, args.reads - n_host, READ_LEN, args.seed + 1 + rank, codes_dev[n_host:], sub_rate=args.sub_rate)
    batch = g.ReadBatch.from_device_bases(codes_dev, READ_LEN, read_id_base=rank * args.reads)
    # raw read bytes in pinned host memory: what the end-to-end path starts from (1 byte/base)
    ascii_lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_host = torch.empty((args.reads, READ_LEN), dtype=torch.uint8, pin_memory=True)
    for a in range(0, args.reads, 5_000_000):
        b = min(args.reads, a + 5_000_000)
        ascii_host[a:b].copy_(ascii_lut[codes_dev[a:b].long()])
    del codes_dev, ref_dev
    torch.cuda.empty_cache()
    caps = dict(mems_per_read=96, recs_per_read=64) if args.random_reads else dict(mems_per_read=24, recs_per_read=8)
    engine = g.Engine(index, args.reads, READ_LEN, **caps)
    lut = g.lut_build(index, LUT_K)
    rmi = None
    if not args.skip_rmi:
        t0 = time.time()
        rmi = train_rmi(index, RMI_K, args.experts, dev)
        log(f"[rank {rank}] RMI probe table + training in {time.time()-t0:.1f}s")
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
    steps_alg = 0
    for a in range(0, n_rec, 50_000_000):
        w = recs[a:a + 50_000_000, 1]
        qs = (w & 0xFFFF).to(torch.int64)
        qe = ((w >> 16) & 0xFFFF).to(torch.int64)
        steps_alg += int(((qe - qs) + (qs > 0).to(torch.int64)).sum().item())
    alg_bytes = 128 * steps_alg + args.reads * ((READ_LEN + 3) // 4) + 16 * n_rec
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (ms_sweep * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tname = {1_000_000_000: "r01c_1gbp_sweep_dram_bytes.json", 100_000_000: "r01_sweep_dram_bytes.json"}.get(args.ref_bases)
    if tname and os.path.exists(os.path.join(ROOT, "profiles", tname)):        # ncu --set full capture of k_sweep on this reference size
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", tname)))
            traffic = tj["dram_bytes_per_read"] * args.reads
            traffic_src = f"profiles/{tname}: {tj['dram_bytes_per_read']:.0f} DRAM bytes/read (ncu dram__bytes_read+write of k_sweep) x reads per launch"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_sweep", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_read": round(alg_bytes / args.reads, 1), "fm_steps_per_read_min": round(steps_alg / args.reads, 2),
                "records_per_read": round(n_rec / args.reads, 3), "ms_sweep": round(ms_sweep, 3), "ms_select_scan_gather": round(ms_sel_bwa, 3),
                "note": "achieved = algorithmic bytes (SURVEY 8d: 128 B per necessary FM step + read + records) / k_sweep time; the access "
                        "pattern is dependent random 64-byte fetches, whose measured ceiling on this GPU is 45.9 G fetches/s = 2.9 TB/s over a "
                        "667 MB index (tools/l2gran_probe.py, profiles/r01_notes.md), not the streaming peak"}

    # ---- the other two methods (device-resident)
    methods = {"bwa": {"reads_per_s": world * args.reads / (ms_bwa * 1e-3), "ms_per_step": ms_bwa}}
    ms_lut = timed(lambda: engine.launch(g.METHOD_LUT, batch, K=LUT_K, lut=lut), max(2, args.steps // 2), 1)
    methods["lut"] = {"reads_per_s": world * args.reads / (ms_lut * 1e-3), "ms_per_step": ms_lut, "K": LUT_K}
    engine.check_overflow()
    if rmi is not None:
        ms_rmi = timed(lambda: engine.launch(g.METHOD_RMI, batch, rmi=rmi), 2, 1)
        methods["rmi"] = {"reads_per_s": world * args.reads / (ms_rmi * 1e-3), "ms_per_step": ms_rmi, "K": RMI_K,
                          "experts": list(args.experts)}
        st = engine.read_status[: args.reads]
        methods["rmi"]["reads_where_reference_raises"] = int((st == g.READ_REF_RAISES).sum().item())
        engine.check_overflow()
    # the device-resident engine's pools are no longer needed: release them before the pipelined path allocates its own
    total_records_dev = engine.records[: n_rec * 16]
    rec_cnt_dev = engine.rec_cnt[: args.reads]

    # ---- end to end through the public API: raw read bytes (pinned host) in, host records out
    pipe = g.PipelinedEngine(index, args.reads, READ_LEN, n_chunks=args.e2e_chunks, **caps)

    def e2e_time(step):
        for _ in range(2):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * args.reads * args.steps / dt, res

    v_ascii, res = e2e_time(lambda: pipe.run_ascii(g.METHOD_BWA, ascii_host, READ_LEN, min_len=1))
    e2e = {"value": v_ascii, "unit": "reads/s", "h2d_bytes_per_step": pipe.last_h2d_bytes, "d2h_bytes_per_step": pipe.last_d2h_bytes,
           "method": "bwa", "records_last_step": int(len(res.records)),
           "api": f"PipelinedEngine.run_ascii ({args.e2e_chunks} chunks, 3 streams: raw ASCII reads H2D | GPU 2-bit packing, sweep, select | "
                  "records D2H, all overlapped)"}
    assert len(res.records) == n_rec, "end-to-end path and device-resident path disagree on the number of records"
    batch.to_host(pin=True)
    v_packed, _ = e2e_time(lambda: pipe.run(g.METHOD_BWA, batch, min_len=1))
    e2e["packed_input"] = {"value": v_packed, "h2d_bytes_per_step": batch.h2d_bytes(),
                           "api": "PipelinedEngine.run (reads already 2-bit packed on the host)"}

    # ---- multi-GPU: the one collective of the path -- gather of per-rank records to rank 0
    total_records = n_rec
    gather = None
    if world > 1:
        from genie_smem_b200 import sharding
        cnt_dev = rec_cnt_dev.to(torch.int64)
        barrier()
        t0 = time.perf_counter()
        g_recs, g_cnts = sharding.gather_records(total_records_dev, cnt_dev, dst=0, device=dev)
        torch.cuda.synchronize()
        dtg = time.perf_counter() - t0
        if rank == 0:
            total_records = int(len(g_recs))
            gather = {"ms": round(dtg * 1e3, 2), "bytes_received": int(len(g_recs)) * 16, "backend": "nccl",
                      "note": "includes the device-to-host read of the gathered array on rank 0; outside the timed steps"}

    # ---- CPU baseline on the host cores, rank 0 at N=1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        text = _B[ref].tobytes()
        sa1 = index.suffix_array_host()
        v, n_s, thr = cpu_arm(text, sa1, reads_head, 0, 0, args.cpu_seconds)
        cpu_baseline = {"value": round(v, 1), "unit": "reads/s", "cores": thr, "kind": "port",
                        "sample": f"first {n_s} reads of the same batch, BWA-SMEM, oracle/smem_oracle.c (reference algorithm, "
                                  f"O(L^2) restarts included) on {thr} pthreads"}
        v2, n2, _ = cpu_arm(text, sa1, reads_head, 1, 0, args.cpu_seconds / 2, K=LUT_K)
        cpu_baseline["lut_reads_per_s"] = round(v2, 1)
        if rmi is not None:
            v3, n3, _ = cpu_arm(text, sa1, reads_head, 2, 0, args.cpu_seconds / 2, rmi=rmi_dict(rmi))
            cpu_baseline["rmi_reads_per_s"] = round(v3, 1)

    if rank == 0:
        out = {"metric": "SMEM reads/sec (BWA-SMEM; LUT/RMI under methods)", "value": world * args.reads / (ms_bwa * 1e-3), "unit": "reads/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_bwa, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": workload,
               "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(gpu_launches),
               "clocks": clocks, "methods": methods, "records_total": total_records, "maximal_matches_per_read": round(n_mems / args.reads, 3),
               "record_gather": gather, "index_build": index.build_stats}
        out["sweep_seed_table_K"] = int(index.c.seed_K)
        emit_result(out)
    if world > 1:
        dist.destroy_process_group()


def train_rmi(index, K, experts, dev, max_keys=8_000_000):
    """RMI over the k-mer -> row keys of this index (reference RMI_LUT.py:36-50).  The {SA value, 32-mer} probe table is
    built on the device (it also serves the last-mile search); a strided sample of its rows is the training set --
    model quality only changes the last-mile length, never the result -- and the fit is the vectorised host trainer."""
    import torch
    import genie_smem_b200 as g
    rmi = g.RmiParams(K, [1] + list(experts), np.zeros(1 + sum(experts)), np.zeros(1 + sum(experts)), dev)
    rmi.build_probe_table(index)
    n_rows = index.n_rows
    stride = max(1, n_rows // max_keys)
    tab = rmi.probe.view(torch.int32).view(-1, 4)[::stride].cpu().numpy().view(np.uint32)
    rows = np.arange(0, n_rows, stride, dtype=np.int64)
    start = tab[:, 0].astype(np.int64) - 1
    code = (tab[:, 1].astype(np.uint64) << np.uint64(32)) | tab[:, 2].astype(np.uint64)
    ok = start + K <= index.n_bases
    key = (code[ok] >> np.uint64(64 - 2 * K)).astype(np.int64)
    m = g.RMI(list(experts)).fit(key, rows[ok])
    out = g.RmiParams(K, m.level_sizes, m.coef, m.intercept, dev)
    out.probe = rmi.probe
    out.c.probe = rmi.probe.data_ptr()
    out.build_none_rows(index)
    return out


def reference_arm(args, workload):
    """--impl reference: the reference's CPU path (oracle port in C, all host threads) on bounded samples of the same
    workload.  The suffix array both arms search is an INPUT of the path: it is built by whichever builder is available
    (the GPU builder when a device is visible, else host SA-IS) before the timed region; the timed search runs on the
    host cores only, through oracle/smem_oracle.c, with none of the engine's kernels."""
    import genie_smem_b200 as g
    ref = make_reference(args.ref_bases, args.seed)
    text = _B[ref].tobytes()
    sa1 = None
    try:
        import torch
        if torch.cuda.is_available():
            idx = g.DeviceIndex.build_on_device(ref, "cuda:0", reverse=False)
            sa1 = idx.suffix_array_host()
            del idx
            torch.cuda.empty_cache()
    except Exception as e:          # no usable device: fall through to the host builder
        log(f"reference arm: GPU index build unavailable ({e}); using host SA-IS")
    if sa1 is None:
        sa1, _ = g.HostIndex.build(text, reverse=False).export()
    reads_codes = host_reads(ref, min(args.reads, HOST_READS), args.seed + 1, args.sub_rate, args.random_reads)
    from oracle.c_oracle import COracle
    o = COracle(text, sa1)
    thr = o.max_threads
    L = READ_LEN
    # size one step to ~8 s
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
    emit_result(out)


if __name__ == "__main__":
    main()
