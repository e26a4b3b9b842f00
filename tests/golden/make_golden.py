#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

This script is the only thing in the repo that touches /root/reference.  It is run once in
the build container (the GPU box has no /root/reference); its outputs are committed.

What it does
  * copies the reference's data files (FASTA + checked-in JSON) into a scratch dir literally
    named .../SMEM (RMI_LUT.py:18-20 needs "SMEM" in the cwd path, ExactMatch.py:32,37,44 use
    cwd-relative "data/"), puts a ~20-line Bio.SeqIO stand-in on sys.path (Biopython is not
    installed; RMI_LUT.py:24-25 only needs record.seq as a str) and imports the five reference
    modules from /root/reference/SMEM unmodified;
  * regenerates the blobs the reference repo is missing (big_data-FM.json, big_data-LUT.json,
    rmi_file.pkl) with the reference's own builders;
  * freezes: index arrays, LUT tables, trained RMI parameters, get_suffix_rmi outputs and the
    SMEM dicts of all three entry points on seeded read sets.

Usage:  python tests/golden/make_golden.py [--scratch /tmp/genie_ref_scratch]
"""
import argparse
import gzip
import hashlib
import json
import os
import random
import shutil
import sys
import time
import warnings

import numpy as np

REF = "/root/reference/SMEM"
OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 20261018

BIO_SHIM = '''\
class _Rec:
    def __init__(self, seq):
        self.seq = seq


def parse(handle, fmt):
    assert fmt == "fasta"
    close = False
    if isinstance(handle, str):
        handle = open(handle, "r")
        close = True
    seq = None
    for line in handle:
        line = line.strip()
        if line.startswith(">"):
            if seq is not None:
                yield _Rec("".join(seq))
            seq = []
        elif seq is not None:
            seq.append(line)
    if seq is not None:
        yield _Rec("".join(seq))
    if close:
        handle.close()
'''


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def setup(scratch):
    smem_dir = os.path.join(scratch, "SMEM")
    data_dir = os.path.join(smem_dir, "data")
    shim_dir = os.path.join(scratch, "shim", "Bio")
    os.makedirs(data_dir, exist_ok=True)
    os.makedirs(shim_dir, exist_ok=True)
    for f in os.listdir(os.path.join(REF, "data")):
        dst = os.path.join(data_dir, f)
        if not os.path.exists(dst) and f != "full_data.fa":
            shutil.copy(os.path.join(REF, "data", f), dst)
            os.chmod(dst, 0o644)
    open(os.path.join(shim_dir, "__init__.py"), "w").close()
    with open(os.path.join(shim_dir, "SeqIO.py"), "w") as f:
        f.write(BIO_SHIM)
    sys.path.insert(0, os.path.join(scratch, "shim"))
    sys.path.insert(0, REF)
    os.chdir(smem_dir)
    return smem_dir


def read_fasta(path):
    with open(path) as f:
        f.readline()
        return "".join(l.strip() for l in f)


def dump_json_gz(name, obj):
    with gzip.open(os.path.join(OUT, name), "wt", compresslevel=9) as f:
        json.dump(obj, f, separators=(",", ":"))


CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def pack2(seq):
    a = np.frombuffer(seq.encode(), dtype=np.uint8)
    lut = np.zeros(256, np.uint8)
    for k, v in CODE.items():
        lut[ord(k)] = v
    c = lut[a]
    pad = (-len(c)) % 4
    c = np.concatenate([c, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def smem_result(fn, *args):
    """Run a reference SMEM entry point; return [[key, lo, hi], ...] in dict order, or
    {"error": <exception class name>} when the reference itself raises."""
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d = fn(*args)
    except (RecursionError, IndexError, TypeError, KeyError) as e:
        return {"error": type(e).__name__}
    out = []
    for k, v in d.items():
        out.append([k, int(v[0]), int(v[1])])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scratch", default="/tmp/genie_ref_scratch")
    args = ap.parse_args()
    sys.setrecursionlimit(1000)  # CPython default; RecursionError is part of the parity domain
    setup(args.scratch)

    from ExactMatch import ExactMatch  # noqa: E402  (reference, unmodified)
    from LUT import LUT  # noqa: E402
    from RMI_LUT import RMI_LUT  # noqa: E402
    import SMEM as SMEM_mod  # noqa: E402
    from SMEM import SMEM  # noqa: E402

    meta = {"seed": SEED, "generated_by": "tests/golden/make_golden.py", "reference": REF}

    # ---------------------------------------------------------------- index arrays
    for name in ("mississippi", "small_data", "medium_data", "big_data"):
        m = ExactMatch(name + ".fa")
        fm_path = os.path.join("data", name + "-FM.json")
        built_here = False
        if not os.path.exists(fm_path):
            t0 = time.time()
            m.create_fm_index()  # n^2 rotations: ~12 s / ~10 GB for big_data
            built_here = True
            print(f"{name}: reference create_fm_index {time.time()-t0:.1f}s")
        else:
            # also prove the checked-in JSON equals what the reference builder produces
            m2 = ExactMatch(name + ".fa")
            m2.load_ref_sequence()
            bwt, first, sa = m2.create_bwt_matrix()
            chk = json.load(open(fm_path))
            assert chk["bwt_array"] == bwt and chk["suffix_array"] == sa
            assert chk["occurance_matrix"] == m2.create_occurance_matrix(bwt)
            assert chk["count_dic"] == m2.create_count_dic(first)
        m.load_fm_index()
        fm = m.fm_index
        text = read_fasta(os.path.join("data", name + ".fa"))
        sa = np.asarray(fm["suffix_array"], dtype=np.uint32)
        bwt = "".join(fm["bwt_array"])
        occ_sha = {c: sha(np.asarray(v, dtype=np.uint32)) for c, v in fm["occurance_matrix"].items()}
        info = {
            "ref_size": fm["ref_size"],
            "count_dic": fm["count_dic"],
            "bwt_sha256": hashlib.sha256(bwt.encode()).hexdigest(),
            "sa_sha256": sha(sa),
            "occ_sha256": occ_sha,
            "built_by_reference_here": built_here,
        }
        if name == "mississippi":
            info["text"] = text
            info["bwt"] = bwt
            info["suffix_array"] = fm["suffix_array"]
            info["occurance_matrix"] = fm["occurance_matrix"]
        else:
            np.savez_compressed(
                os.path.join(OUT, f"index_{name}.npz"),
                text2bit=pack2(text),
                n_bases=np.int64(len(text)),
                suffix_array=sa,
                bwt=np.frombuffer(bwt.encode(), dtype=np.uint8),
            )
        meta[name] = info

    # ---------------------------------------------------------------- mississippi known answers
    m = ExactMatch("mississippi.fa")
    m.load_fm_index()
    s = SMEM.__new__(SMEM)  # SMEM.__init__ would load the legacy-format mississippi LUT
    s.matcher = m
    known = {
        "exact_match_back_prop": {q: m.exact_match_back_prop(q) for q in ["iss", "ssi", "i", "p", "mississippi", "", "ssip"]},
        "exact_match": {q: m.exact_match(q) for q in ["iss", "ssi", "i"]},
        "miss": {q: m.exact_match_back_prop(q) for q in ["sm", "ipi", "pm"]},
        "get_SMEMS": {},
    }
    for q, ml in [("pissssi", 1), ("mmissippss", 1), ("mississippi", 1), ("ssissim", 2), ("ipsmipsi", 1)]:
        known["get_SMEMS"][f"{q}|{ml}"] = smem_result(s.get_SMEMS, q, ml)
    meta["mississippi"]["known"] = known

    # ---------------------------------------------------------------- LUT tables
    # medium K=6: the reference's own checked-in JSON
    lj = json.load(open(os.path.join("data", "medium_data-LUT.json")))
    keys = sorted(int(k) for k in lj["lut"])
    np.savez_compressed(
        os.path.join(OUT, "lut_medium_data_k6.npz"),
        K=np.int64(lj["lut_size"]),
        keys=np.asarray(keys, np.uint32),
        lo=np.asarray([lj["lut"][str(k)][0][0] for k in keys], np.uint32),
        hi=np.asarray([lj["lut"][str(k)][0][1] for k in keys], np.uint32),
        npos=np.asarray([len(lj["lut"][str(k)][1]) for k in keys], np.uint32),
        pos=np.asarray([p for k in keys for p in lj["lut"][str(k)][1]], np.uint32),
    )
    # big K=12: regenerate with the reference builder (8 s)
    big_lut_path = os.path.join("data", "big_data-LUT.json")
    if not os.path.exists(big_lut_path):
        mb = ExactMatch("big_data.fa")
        mb.load_fm_index()
        lb = LUT(mb)
        t0 = time.time()
        lb.generate_lut(12)
        lb.save_lut()
        print(f"big_data: reference generate_lut(12) {time.time()-t0:.1f}s, {len(lb.lut)} entries")
    lj = json.load(open(big_lut_path))
    keys = sorted(int(k) for k in lj["lut"])
    np.savez_compressed(
        os.path.join(OUT, "lut_big_data_k12.npz"),
        K=np.int64(lj["lut_size"]),
        keys=np.asarray(keys, np.uint32),
        lo=np.asarray([lj["lut"][str(k)][0][0] for k in keys], np.uint32),
        hi=np.asarray([lj["lut"][str(k)][0][1] for k in keys], np.uint32),
        pos_sha256=np.asarray(sha(np.asarray([p for k in keys for p in lj["lut"][str(k)][1]], np.uint32))),
    )

    # ---------------------------------------------------------------- RMI models + lookups
    rng = random.Random(SEED)
    rmis = {}
    for ref_name, K, experts in [("medium_data", 6, [10, 100]), ("big_data", 12, [10, 100]), ("big_data", 15, [10, 100])]:
        tag = f"{ref_name}_k{K}"
        r = RMI_LUT(list(experts), K, ref_name + ".fa")
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            r.train_RMI()
        rmis[tag] = r
        levels = [len(l) for l in r.rmi.models]
        coef = np.asarray([float(mo.coef_[0]) for lvl in r.rmi.models for mo in lvl], np.float64)
        icpt = np.asarray([float(mo.intercept_) for lvl in r.rmi.models for mo in lvl], np.float64)
        # golden lookups: present k-mers (sampled from the text) + uniform random k-mers
        text = read_fasta(os.path.join("data", ref_name + ".fa"))
        qs = []
        for _ in range(1500):
            p = rng.randrange(0, len(text) - K + 1)
            qs.append(text[p:p + K])
        for _ in range(1500):
            qs.append("".join(rng.choice("ACGT") for _ in range(K)))
        qs += ["A" * K, "C" * K, "G" * K, "T" * K]
        # the K-mers hanging over the end of the text exercise the short-suffix (None) rows
        qs += [text[-K:], text[-K - 1:-1], text[:K]]
        res = []
        for q in qs:
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    pred = float(r.rmi_predict(q)[0])
                    lo, hi = r.get_suffix_rmi(q)
                res.append([q, pred, int(lo), int(hi)])
            except (RecursionError, IndexError) as e:
                res.append([q, None, type(e).__name__, 0])
        np.savez_compressed(os.path.join(OUT, f"rmi_{tag}.npz"), K=np.int64(K), experts=np.asarray(experts, np.int64),
                            level_sizes=np.asarray(levels, np.int64), coef=coef, intercept=icpt)
        dump_json_gz(f"rmi_lookups_{tag}.json.gz", res)
        print(f"rmi {tag}: levels {levels}, {sum(1 for x in res if x[1] is None)} lookups raised")

    # ---------------------------------------------------------------- SMEM golden sets
    def hoist(r):
        RMI_LUT.load = staticmethod(lambda file, _r=r: _r)  # SMEM.py:207 reloads per call; hoist

    def make_reads(text, n_exact, n_random, n_pieces, L, rng):
        reads = []
        for _ in range(n_exact):
            p = rng.randrange(0, len(text) - L + 1)
            reads.append(text[p:p + L])
        for _ in range(n_random):
            reads.append("".join(rng.choice("ACGT") for _ in range(L)))
        for _ in range(n_pieces):  # create_query_from_ref (SMEM.py:496-505) on the text without '$'
            q = ""
            while len(q) < L:
                pos = rng.randint(0, len(text))
                size = rng.randint(1, 30)
                if size + pos > len(text):
                    continue
                q += text[pos:pos + size]
            reads.append(q[:L])
        return reads

    def run_set(tag, ref_name, reads, K_lut, rmi_tags, minlens=(1,)):
        m = ExactMatch(ref_name + ".fa")
        m.load_fm_index()
        s = SMEM(m)  # loads <ref>-LUT.json
        assert s.lut.lut_size == K_lut
        out = {"ref": ref_name, "K_lut": K_lut, "reads": reads, "bwa": {}, "lut": None, "rmi": {}}
        t0 = time.time()
        for ml in minlens:
            out["bwa"][str(ml)] = [smem_result(s.get_SMEMS, q, ml) for q in reads]
        t1 = time.time()
        out["lut"] = [smem_result(s.get_smems_lut, q) if len(q) >= K_lut else None for q in reads]
        t2 = time.time()
        for rt in rmi_tags:
            hoist(rmis[rt])
            K = rmis[rt].prediction_size
            out["rmi"][rt] = [smem_result(s.get_smems_rmi, q) if len(q) >= K else None for q in reads]
        t3 = time.time()
        out["ref_seconds"] = {"bwa": t1 - t0, "lut": t2 - t1, "rmi": t3 - t2}
        dump_json_gz(f"smems_{tag}.json.gz", out)
        nerr = sum(1 for rt in rmi_tags for x in out["rmi"][rt] if isinstance(x, dict))
        print(f"smems {tag}: {len(reads)} reads, ref seconds {out['ref_seconds']}, rmi raised on {nerr}")

    big = read_fasta(os.path.join("data", "big_data.fa"))
    med = read_fasta(os.path.join("data", "medium_data.fa"))
    rng = random.Random(SEED)
    # C1: 1,000 exact 101-bp reads (BASELINE.json configs[0])
    run_set("c1_big_exact101", "big_data", make_reads(big, 1000, 0, 0, 101, rng), 12, ["big_data_k15"], minlens=(1, 20))
    # C2: 500 exact + 500 random (+ 300 reference-piece reads), LUT K=12, RMI K=15 and K=12 (configs[1])
    run_set("c2_big_mixed101", "big_data", make_reads(big, 500, 500, 300, 101, rng), 12, ["big_data_k15", "big_data_k12"], minlens=(1, 12))
    # 151-bp reads with 1 % substitutions (the C3-C5 read model, on the small reference)
    reads = []
    for q in make_reads(big, 300, 0, 0, 151, rng):
        q = list(q)
        for i in range(len(q)):
            if rng.random() < 0.01:
                q[i] = rng.choice([c for c in "ACGT" if c != q[i]])
        reads.append("".join(q))
    run_set("big_sub151", "big_data", reads, 12, ["big_data_k15"])
    # fuzz on medium_data (K=6): ragged lengths 6..151, all three read kinds
    reads = []
    for _ in range(1500):
        L = rng.randint(6, 151)
        kind = rng.randrange(3)
        reads += make_reads(med, int(kind == 0), int(kind == 1), int(kind == 2), L, rng)
    run_set("medium_fuzz", "medium_data", reads, 6, ["medium_data_k6"], minlens=(1, 8))

    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("done")


if __name__ == "__main__":
    main()
