#!/usr/bin/env python
"""Exhaustive small-world check of the identity gsm_smem_select(LUT) relies on:

    get_smems_lut(q) == get_SMEMS(q, min_len = 1), record for record, for every read q of at least K bases.

Both sides are the C restatement of the reference (oracle/smem_oracle.c, pinned to tests/golden): every reference text over
ACGT of length 4..n_max that contains all four bases x every read over ACGT of length 1..l_max x K = 1..k_max.  Test
infrastructure only (imports oracle/; lives under tests/ for that reason); tests/test_oracle_c.py runs a bounded slice of it.

    python tests/offline/lut_identity_exhaustive.py [n_max=6] [l_max=7] [k_max=4]
    python tests/offline/lut_identity_exhaustive.py ac [n_max=12] [l_max=10] [k_max=6]   # repetitive worlds: references = every string
                                                                                 # over AC of 2..n_max bases + "GT", reads over AC
"""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def all_reads(l_max, alphabet="ACGT"):
    reads = []
    for L in range(1, l_max + 1):
        reads += ["".join(t) for t in itertools.product(alphabet, repeat=L)]
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


def texts_ac(n_min, n_max):
    for n in range(n_min, n_max + 1):
        for t in itertools.product("AC", repeat=n):
            if len(set(t)) == 2:                 # all four bases present, as above
                yield "".join(t) + "GT"


def _work(args):
    text, l_max, k_max, alphabet = args
    reads = _work.reads if getattr(_work, "key", None) == (l_max, alphabet) else None
    if reads is None:
        _work.reads, _work.key = all_reads(l_max, alphabet), (l_max, alphabet)
        _work.lens = np.asarray([len(r) for r in _work.reads], np.uint32)
        _work.joined = "".join(_work.reads).encode()
    return check_text(text, _work.reads, _work.joined, _work.lens, k_max)


def main():
    argv = sys.argv[1:]
    ac = bool(argv) and argv[0] == "ac"
    if ac:
        argv = argv[1:]
    n_max = int(argv[0]) if len(argv) > 0 else (12 if ac else 6)
    l_max = int(argv[1]) if len(argv) > 1 else (10 if ac else 7)
    k_max = int(argv[2]) if len(argv) > 2 else (6 if ac else 4)
    from multiprocessing import Pool
    jobs = [(t, l_max, k_max, "AC" if ac else "ACGT") for t in (texts_ac(2, n_max) if ac else texts(4, n_max))]
    with Pool(os.cpu_count()) as pool:
        total = sum(pool.imap_unordered(_work, jobs, chunksize=8))
    world = (f"every string over AC of 2..{n_max} bases + 'GT', reads over AC" if ac else
             f"length 4..{n_max} over ACGT, all four bases present, reads over ACGT")
    print(f"{len(jobs)} references ({world}) x every read of length 1..{l_max} x K = 1..{k_max}: "
          f"{total} cases, get_smems_lut == get_SMEMS(min_len 1) in every one")


if __name__ == "__main__":
    main()
