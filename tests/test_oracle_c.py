"""Pin oracle/smem_oracle.c (the C restatement used as CPU baseline and bulk checker) against the
golden fixtures frozen from the live reference."""
import numpy as np
import pytest

from oracle.c_oracle import COracle
from tests import golden_util as gu


@pytest.fixture(scope="module")
def orcs():
    out = {}
    for name in ("small_data", "medium_data", "big_data"):
        g = gu.load_index(name)
        out[name] = (g, COracle(g["text"], g["suffix_array"]))
    return out


def test_backsearch_matches_python_port(orcs):
    from oracle import ref_port as rp
    g, o = orcs["medium_data"]
    idx = rp.RefIndex(g["text"], g["suffix_array"])
    rng = np.random.default_rng(3)
    reads = []
    for _ in range(300):
        L = int(rng.integers(1, 60))
        p = int(rng.integers(0, len(g["text"]) - L))
        reads.append(g["text"][p:p + L])
        reads.append("".join("ACGT"[c] for c in rng.integers(0, 4, L)))
    lo, hi = o.backsearch(reads)
    for q, l, h in zip(reads, lo, hi):
        e = idx.exact_match_back_prop(q)
        assert (-1 if h < l else (int(l), int(h))) == e


@pytest.mark.parametrize("tag,name", [("medium_data_k6", "medium_data"), ("big_data_k12", "big_data"), ("big_data_k15", "big_data")])
def test_rmi_lookups(orcs, tag, name):
    _, o = orcs[name]
    p = gu.load_rmi(tag)
    for q, pred, lo, hi in gu.load_json(f"rmi_lookups_{tag}.json.gz"):
        code = 0
        for ch in q:
            code = code << 2 | "ACGT".index(ch)
        st, gp, glo, ghi = o.rmi_lookup(p, code)
        if pred is None:
            assert st == -1
        else:
            assert st == 0 and gp == pred and (glo, ghi) == (lo, hi)


@pytest.mark.parametrize("fname", ["smems_c1_big_exact101.json.gz", "smems_c2_big_mixed101.json.gz", "smems_big_sub151.json.gz",
                                   "smems_medium_fuzz.json.gz"])
def test_smem_sets(orcs, fname):
    g = gu.load_json(fname)
    _, o = orcs[g["ref"]]
    reads = g["reads"]
    for ml, exp in g["bwa"].items():
        assert o.smem_dicts(0, reads, min_len=int(ml)) == exp
    sel = [i for i, e in enumerate(g["lut"]) if e is not None]
    assert o.smem_dicts(1, [reads[i] for i in sel], K=g["K_lut"]) == [g["lut"][i] for i in sel]
    for tag, exp in g["rmi"].items():
        p = gu.load_rmi(tag)
        sel = [i for i, e in enumerate(exp) if e is not None]
        got = o.smem_dicts(2, [reads[i] for i in sel], rmi=p)
        for k, i in enumerate(sel):
            if isinstance(exp[i], dict):
                assert got[k] == "raises"
            else:
                assert got[k] == exp[i]


def test_lut_identity_exhaustive_slice():
    """get_smems_lut == get_SMEMS(min_len 1) record for record on reads of at least K bases -- the identity
    gsm_smem_select(LUT) relies on -- exhaustively on a small world: every reference over ACGT of 4..5
    bases that contains all four, every read of 1..5 bases, K = 1..3 (tests/offline/lut_identity_exhaustive.py ran 4..7 x 1..7 x 1..4)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "lut_identity_exhaustive", os.path.join(os.path.dirname(__file__), "offline", "lut_identity_exhaustive.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    reads = tool.all_reads(5)
    lens = np.asarray([len(r) for r in reads], np.uint32)
    joined = "".join(reads).encode()
    n = sum(tool.check_text(t, reads, joined, lens, 3) for t in tool.texts(4, 5))
    assert n > 1_000_000


def test_base_absent_from_the_reference_raises():
    """count_dic[char] raises KeyError for a base the reference does not contain (ExactMatch.py:140): the oracle reports the
    read as 'raises' (and terminates) instead of walking a position it can never pass."""
    import genie_smem_b200 as gs
    text = "AAGTAGGTTA"
    sa, _ = gs.HostIndex.build(text).export()
    o = COracle(text, sa)
    reads = ["AAGT", "ACGT", "C", "GGTT", "TTAC"]
    for method, kw in ((0, {"min_len": 1}), (1, {"K": 2})):
        got = o.smem_dicts(method, reads, **kw)
        assert [g == "raises" for g in got] == [False, True, True, False, True]
        assert got[0][0][0] == "AAGT" and got[3][0][0] == "GGTT"


def test_rmi_identity_exhaustive_slice():
    """get_smems_rmi == get_SMEMS(min_len 1) for every read whose windows all look up exactly -- the identity behind the RMI-SMEM
    pre-filter -- exhaustively on a small world: every reference over ACGT of 4..5 bases with all four
    bases, a trained and a perturbed model, K = 1..3, every read of K..5 bases (tests/offline/rmi_identity_exhaustive.py ran
    larger worlds)."""
    import importlib.util
    import itertools
    import os
    spec = importlib.util.spec_from_file_location(
        "rmi_identity_exhaustive", os.path.join(os.path.dirname(__file__), "offline", "rmi_identity_exhaustive.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    texts = ["".join(t) for n in (4, 5) for t in itertools.product("ACGT", repeat=n) if len(set(t)) == 4]
    n_same = n_skip = 0
    for i, t in enumerate(texts[::3]):
        a, b = tool.check_text(t, 5, 3, i)
        n_same += a
        n_skip += b
    assert n_same > 50_000 and n_skip > 0
