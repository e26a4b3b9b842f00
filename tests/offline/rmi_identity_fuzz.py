#!/usr/bin/env python
"""Fuzz of the identity behind the RMI-SMEM pre-filter on the literal Python restatement of the reference
(oracle/ref_port.py): on random, periodic and palindromic small references with trained and perturbed two-level models, every read
whose K-mer windows ALL look up exactly through get_suffix_rmi (hit <=> the k-mer occurs, true interval) must satisfy
get_smems_rmi(q) == get_SMEMS(q, 1) as ordered dicts.  Test infrastructure (imports oracle/).

    python tests/offline/rmi_identity_fuzz.py [seed] [seconds]        (round 2: 6 seeds x 420 s = 4.79 M identical cases, 0 differences)
"""
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import genie_smem_b200 as gs
from oracle import ref_port as rp
from __graft_entry__ import _rmi_keys
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 60
n_same = n_skip = n_worlds = 0
while time.time() < t_end:
    n_ref = rnd.choice((60, 120, 300, 800))
    mode = rnd.random()
    alpha = rnd.choice(("ACGT", "ACGT", "AC", "AAAC", "ACG"))
    if mode < 0.4:
        text = "".join(rnd.choice(alpha) for _ in range(n_ref))
    elif mode < 0.8:
        unit = "".join(rnd.choice("ACGT") for _ in range(rnd.choice((2, 3, 5, 7, 11))))
        t = list((unit * (n_ref // len(unit) + 1))[:n_ref])
        pm = rnd.choice((0.0, 0.03, 0.1))
        for k in range(len(t)):
            if rnd.random() < pm:
                t[k] = rnd.choice("ACGT")
        text = "".join(t)
    else:
        a = "".join(rnd.choice("ACGT") for _ in range(n_ref // 3))
        text = a + a[::-1] + a
    if len(set(text)) < 4:
        text += rnd.choice(("ACGT", "TGCA", "GATC"))
    sa, _ = gs.HostIndex.build(text).export()
    idx = rp.RefIndex(text, sa)
    K = rnd.choice((2, 3, 4, 5, 6))
    try:
        m = gs.RMI([rnd.choice((1, 2, 4, 8))]).fit(*_rmi_keys(text, sa, K))
    except Exception as e:
        continue
    icpt = np.array(m.intercept, np.float64)
    if rnd.random() < 0.5:
        icpt[1:] += np.random.default_rng(rnd.randrange(1 << 30)).integers(-6, 6, len(icpt) - 1)
    rmi = rp.RefRMI(idx, K, m.level_sizes, m.coef, icpt)
    o = rp.RefSMEM(idx, rmi=rmi)
    n_worlds += 1
    cache = {}
    for _ in range(60):
        L = rnd.choice((K, K + 1, 2 * K, 12, 25, 50))
        r = rnd.random()
        if r < 0.6:
            L = min(L, len(text) - 1)
            s0 = rnd.randrange(0, max(1, len(text) - L))
            q = list(text[s0:s0 + L])
            pm = rnd.choice((0.0, 0.03, 0.1, 0.3))
            for k in range(len(q)):
                if rnd.random() < pm:
                    q[k] = rnd.choice("ACGT")
            q = "".join(q)
        else:
            q = "".join(rnd.choice(alpha if r < 0.8 else "ACGT") for _ in range(L))
        if len(q) < K:
            continue
        exact = True
        for i in range(len(q) - K + 1):
            kmer = q[i:i + K]
            if kmer not in cache:
                true = idx.exact_match_back_prop(kmer)
                try:
                    got = rmi.get_suffix_rmi(kmer)
                    cache[kmer] = not ((true == -1 and got[1] >= got[0]) or (true != -1 and tuple(got) != tuple(true)))
                except (IndexError, RecursionError, TypeError):
                    cache[kmer] = False
            if not cache[kmer]:
                exact = False
                break
        if not exact:
            n_skip += 1
            continue
        a = o.get_SMEMS(q, 1)
        b = {k: tuple(v) for k, v in o.get_smems_rmi(q).items()}
        assert a == b and list(a) == list(b), (text, K, q, list(m.level_sizes), list(m.coef), list(icpt))
        n_same += 1
print("worlds", n_worlds, "identical", n_same, "skipped (a window not exact)", n_skip)
