"""Loaders for the fixtures under tests/golden/ (produced by tests/golden/make_golden.py)."""
import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_B = np.frombuffer(b"ACGT", dtype=np.uint8)


def meta():
    with open(os.path.join(GOLDEN, "meta.json")) as f:
        return json.load(f)


def unpack2(packed, n):
    p = np.asarray(packed, dtype=np.uint8)
    c = np.stack([p & 3, (p >> 2) & 3, (p >> 4) & 3, (p >> 6) & 3], axis=1).reshape(-1)[:n]
    return _B[c].tobytes().decode()


def load_index(name):
    """-> dict(text=str, suffix_array=np.uint32[n], bwt=str)"""
    if name == "mississippi":
        m = meta()["mississippi"]
        return {"text": m["text"], "suffix_array": np.asarray(m["suffix_array"], np.uint32), "bwt": m["bwt"]}
    z = np.load(os.path.join(GOLDEN, f"index_{name}.npz"))
    return {"text": unpack2(z["text2bit"], int(z["n_bases"])), "suffix_array": z["suffix_array"],
            "bwt": z["bwt"].tobytes().decode()}


def load_lut(tag):
    return dict(np.load(os.path.join(GOLDEN, f"lut_{tag}.npz")))


def load_rmi(tag):
    z = np.load(os.path.join(GOLDEN, f"rmi_{tag}.npz"))
    return {"K": int(z["K"]), "experts": [int(x) for x in z["experts"]], "level_sizes": [int(x) for x in z["level_sizes"]],
            "coef": z["coef"], "intercept": z["intercept"]}


def load_json(name):
    with gzip.open(os.path.join(GOLDEN, name), "rt") as f:
        return json.load(f)


def norm(d):
    """reference dict -> [[key, lo, hi], ...] in insertion order (list/tuple artefact removed)"""
    return [[k, int(v[0]), int(v[1])] for k, v in d.items()]
