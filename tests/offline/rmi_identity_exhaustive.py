#!/usr/bin/env python
"""Exhaustive small-world check of the identity behind the RMI-SMEM pre-filter, on the C restatement of
the reference (oracle/smem_oracle.c, pinned to tests/golden):

    every K-mer window of q looks up exactly through get_suffix_rmi  ==>  get_smems_rmi(q) == get_SMEMS(q, min_len = 1)

for every reference over ACGT of 4..n_max bases that contains all four bases x a trained two-level model and a perturbed one x
K = 1..k_max x EVERY read of K..l_max bases.  "Exactly" = the lookup does not raise, hits iff the k-mer occurs and then returns
its true interval (orc_rmi_lookup against orc_backsearch, per code).  Test infrastructure (imports oracle/).

    python tests/offline/rmi_identity_exhaustive.py [n_max=6] [l_max=7] [k_max=3]

Round 2: n_max 7, l_max 7, k_max 3: 209 M reads with exact windows; n_max 8, l_max 7, k_max 4: 1,712 M; no difference.
"""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def _codes_of(reads_arr, K):
    """window codes of equal-length reads given as (n, L) arrays of base codes -> (n, L-K+1)"""
    n, L = reads_arr.shape
    out = np.zeros((n, L - K + 1), np.int64)
    for t in range(K):
        out = out * 4 + reads_arr[:, t:L - K + 1 + t]
    return out


def check_text(text, l_max, k_max, seed):
    import genie_smem_b200 as gs
    from oracle.c_oracle import COracle
    from __graft_entry__ import _rmi_keys
    sa, _ = gs.HostIndex.build(text).export()
    o = COracle(text, sa)
    rng = np.random.default_rng(seed)
    n_same = n_skip = 0
    for K in range(1, min(k_max, len(text) - 1) + 1):
        keys, rows = _rmi_keys(text, sa, K)
        if len(keys) < 2:
            continue
        m = gs.RMI([int(rng.choice((1, 2, 4)))]).fit(keys, rows)
        kmers = ["".join("ACGT"[(c >> (2 * (K - 1 - t))) & 3] for t in range(K)) for c in range(4 ** K)]
        tlo, thi = o.backsearch(kmers)
        for perturbed in (False, True):
            icpt = np.array(m.intercept, np.float64)
            if perturbed:
                icpt[1:] += rng.integers(-3, 4, len(icpt) - 1)
            rmi = {"K": K, "level_sizes": list(m.level_sizes), "coef": np.asarray(m.coef, np.float64), "intercept": icpt}
            exact = np.zeros(4 ** K, bool)
            for c in range(4 ** K):
                st, _, lo, hi = o.rmi_lookup(rmi, c)
                occurs = thi[c] >= tlo[c]
                exact[c] = st == 0 and ((not occurs and hi < lo) or (occurs and (lo, hi) == (int(tlo[c]), int(thi[c]))))
            for L in range(K, l_max + 1):
                arr = np.array(list(itertools.product(range(4), repeat=L)), np.int64)
                ok = exact[_codes_of(arr, K)].all(axis=1)
                n_skip += int((~ok).sum())
                if not ok.any():
                    continue
                sel = arr[ok]
                joined = np.frombuffer(b"ACGT", np.uint8)[sel.reshape(-1)].tobytes()
                lens = np.full(len(sel), L, np.uint32)
                a, an = o.smems(0, None, min_len=1, joined=joined, lens=lens, threads=1)
                b, bn = o.smems(2, None, rmi=rmi, joined=joined, lens=lens, threads=1)
                same = (an == bn) & (a == b).all(axis=(1, 2))
                if not same.all():
                    k = int(np.nonzero(~same)[0][0])
                    raise AssertionError((text, K, perturbed, "".join("ACGT"[x] for x in sel[k]), rmi))
                n_same += len(sel)
    return n_same, n_skip


def _work(args):
    return check_text(*args)


def main():
    argv = sys.argv[1:]
    n_max = int(argv[0]) if len(argv) > 0 else 6
    l_max = int(argv[1]) if len(argv) > 1 else 7
    k_max = int(argv[2]) if len(argv) > 2 else 3
    from multiprocessing import Pool
    texts = ["".join(t) for n in range(4, n_max + 1) for t in itertools.product("ACGT", repeat=n) if len(set(t)) == 4]
    with Pool(os.cpu_count()) as pool:
        res = list(pool.imap_unordered(_work, [(t, l_max, k_max, i) for i, t in enumerate(texts)], chunksize=4))
    print(f"{len(texts)} references (length 4..{n_max}, all four bases) x 2 models x K = 1..{k_max} x every read of K..{l_max} bases: "
          f"{sum(r[0] for r in res)} reads with exact windows, get_smems_rmi == get_SMEMS(min_len 1) on every one; "
          f"{sum(r[1] for r in res)} reads skipped (a window does not look up exactly)")


if __name__ == "__main__":
    main()
