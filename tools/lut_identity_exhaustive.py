#!/usr/bin/env python
"""Exhaustive small-world check of the identity gsm_smem_select(LUT) relies on:

    get_smems_lut(q) == get_SMEMS(q, min_len = 1), record for record, for every read q of at least K bases.

Both sides are the C restatement of the reference (oracle/smem_oracle.c, pinned to tests/golden): every reference text over
ACGT of length 4..n_max that contains all four bases x every read over ACGT of length 1..l_max x K = 1..k_max.  Test
infrastructure only (imports oracle/); tests/test_logic_emu.py runs a bounded slice of it.

    python tools/lut_identity_exhaustive.py [n_max=6] [l_max=7] [k_max=4]
"""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def all_reads(l_max):
    reads = []
    for L in range(1, l_max + 1):
        reads += ["".join(t) for t in itertools.product("ACGT", repeat=L)]
    return reads


def check_text(text, reads, joined, lens, k_max):
    """Number of (read, K) cases compared on this text; raises AssertionError on the first difference."""
    import genie_smem_b200 as gs
    from oracle.c_oracle import COracle
    sa, _ = gs.HostIndex.build(text).export()
    o = COracle(text, sa)
    bwa, bwa_n = o.smems(0, None, min_len=1, joined=joined, lens=lens, threads=1)
    n = 0
    for K in range(1, k_max + 1):
        lut, lut_n = o.smems(1, None, K=K, joined=joined, lens=lens, threads=1)
        long_enough = lens >= K
        assert (lut_n[~long_enough] == -2).all(), (text, K, "reads shorter than K must be flagged")
        same = (lut_n == bwa_n) & (lut == bwa).all(axis=(1, 2))
        bad = np.nonzero(long_enough & ~same)[0]
        assert bad.size == 0, (text, K, reads[int(bad[0])])
        n += int(long_enough.sum())
    return n


def texts(n_min, n_max):
    for n in range(n_min, n_max + 1):
        for t in itertools.product("ACGT", repeat=n):
            if len(set(t)) == 4:                 # the parity domain: references containing all four bases
                yield "".join(t)


def _work(args):
    text, l_max, k_max = args
    reads = _work.reads if getattr(_work, "l_max", None) == l_max else None
    if reads is None:
        _work.reads, _work.l_max = all_reads(l_max), l_max
        _work.lens = np.asarray([len(r) for r in _work.reads], np.uint32)
        _work.joined = "".join(_work.reads).encode()
    return check_text(text, _work.reads, _work.joined, _work.lens, k_max)


def main():
    n_max = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    l_max = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    k_max = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    from multiprocessing import Pool
    jobs = [(t, l_max, k_max) for t in texts(4, n_max)]
    with Pool(os.cpu_count()) as pool:
        total = sum(pool.imap_unordered(_work, jobs, chunksize=8))
    print(f"{len(jobs)} references (length 4..{n_max}, all four bases) x every read of length 1..{l_max} x K = 1..{k_max}: "
          f"{total} cases, get_smems_lut == get_SMEMS(min_len 1) in every one")


if __name__ == "__main__":
    main()
