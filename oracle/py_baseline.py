"""TEST / BENCH INFRASTRUCTURE (never imported by the product): the literal Python restatement of the reference
(oracle/ref_port.py, pinned to tests/golden) timed on the host cores, as BASELINE.md section 3 asks for the reference's own
Python -- one process (reads/s/core) and multiprocessing.Pool(os.cpu_count()) with the core count stated -- for get_SMEMS,
get_smems_lut (K = 12) and get_smems_rmi (golden [10, 100] model, K = 15; the model load of SMEM.py:207 hoisted out of the
call) on BASELINE.json configs[0]: big_data (100 kb), 101-bp exact substrings.  /root/reference does not travel to the GPU
box, which is why the restatement stands in for it.  Runs in its own process (bench.py starts it with subprocess) so that
the pool forks from a process that never touched CUDA.

    python -m oracle.py_baseline [--seconds S]      ->  one JSON line
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_S = None


def _smem():
    global _S
    if _S is None:
        from oracle import ref_port as rp
        from tests import golden_util as gu
        g = gu.load_index("big_data")
        idx = rp.RefIndex(g["text"], g["suffix_array"])
        p = gu.load_rmi("big_data_k15")
        _S = (g["text"], rp.RefSMEM(idx, lut=rp.RefLUT(idx, 12), rmi=rp.RefRMI(idx, p["K"], p["level_sizes"], p["coef"], p["intercept"])))
    return _S


def _reads(n, seed):
    text, _ = _smem()
    rng = np.random.default_rng(seed)
    return [text[p:p + 101] for p in rng.integers(0, len(text) - 101, n)]


def _run(method, reads):
    _, s = _smem()
    n_raise = 0
    for q in reads:
        if method == "bwa":
            s.get_SMEMS(q, 1)
        elif method == "lut":
            s.get_smems_lut(q)
        else:
            try:
                s.get_smems_rmi(q)
            except (IndexError, RecursionError, TypeError):      # the reference raises on some reads
                n_raise += 1
    return n_raise


def _pool_job(args):
    method, n, seed = args
    _run(method, _reads(n, seed))
    return n


def one_process(method, seconds):
    reads = _reads(4000, 20261018)
    _run(method, reads[:3])
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds and n < len(reads):
        _run(method, reads[n:n + 10])
        n += 10
    return n / (time.perf_counter() - t0), n


def pooled(method, seconds, per_job=25):
    from multiprocessing import Pool
    cores = os.cpu_count() or 1
    with Pool(cores, initializer=_smem) as pool:
        pool.map(_pool_job, [(method, 2, k) for k in range(cores)])            # every worker has built its index
        rate1, _ = one_process(method, min(1.0, seconds))
        jobs = max(cores, int(rate1 * cores * seconds / per_job))
        t0 = time.perf_counter()
        done = sum(pool.imap_unordered(_pool_job, [(method, per_job, 1000 + k) for k in range(jobs)]))
        dt = time.perf_counter() - t0
    return done / dt, done, cores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0, help="budget per measurement")
    a = ap.parse_args()
    out = {"what": "oracle/ref_port.py (literal Python restatement of SMEM.py get_SMEMS / get_smems_lut / get_smems_rmi with the model "
                   "load hoisted), BASELINE.json configs[0]: big_data (100 kb), 101-bp exact substrings, LUT K = 12, RMI [10, 100] K = 15",
           "cores": os.cpu_count() or 1}
    for m in ("bwa", "lut", "rmi"):
        r, n = one_process(m, a.seconds)
        out[f"{m}_reads_per_s_one_process"] = round(r, 1)
        out[f"{m}_reads_one_process"] = n
    for m in ("bwa", "lut", "rmi"):
        r, n, cores = pooled(m, a.seconds)
        out[f"{m}_reads_per_s_pool"] = round(r, 1)
        out[f"{m}_reads_pool"] = n
    out["reads_per_s_one_process"] = out["bwa_reads_per_s_one_process"]         # the key earlier records carry
    out["reads"] = out["bwa_reads_one_process"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
