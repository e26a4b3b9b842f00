"""Pin oracle/ref_port.py (the CPU restatement) against fixtures frozen from the live reference."""
import hashlib

import numpy as np
import pytest

from oracle import ref_port as rp
from tests import golden_util as gu


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def indexes():
    out = {}
    for name in ("mississippi", "small_data", "medium_data", "big_data"):
        g = gu.load_index(name)
        out[name] = (g, rp.RefIndex(g["text"]))
    return out


@pytest.mark.parametrize("name", ["mississippi", "small_data", "medium_data", "big_data"])
def test_index_arrays_equal_reference(indexes, name):
    g, idx = indexes[name]
    m = gu.meta()[name]
    assert idx.ref_size == m["ref_size"]
    assert np.array_equal(idx.suffix_array, g["suffix_array"].astype(np.int64))
    assert idx.bwt == g["bwt"]
    assert idx.count_dic == m["count_dic"]
    for c, h in m["occ_sha256"].items():
        assert _sha(idx.occ[c].astype(np.uint32)) == h


def test_mississippi_known_answers(indexes):
    _, idx = indexes["mississippi"]
    known = gu.meta()["mississippi"]["known"]
    assert idx.exact_match_back_prop("iss") == (3, 4)      # paper Fig. 2
    assert idx.exact_match("iss") == [2, 5]
    for q, v in known["exact_match_back_prop"].items():
        assert list(idx.exact_match_back_prop(q)) == v
    for q, v in known["miss"].items():
        assert idx.exact_match_back_prop(q) == v == -1
    for q, v in known["exact_match"].items():
        assert idx.exact_match(q) == v
    s = rp.RefSMEM(idx)
    for key, v in known["get_SMEMS"].items():
        q, ml = key.split("|")
        assert gu.norm(s.get_SMEMS(q, int(ml))) == v
    assert dict(s.get_SMEMS("pissssi", 1)) == {"pi": (6, 6), "iss": (3, 4), "ssi": (10, 11)}


@pytest.mark.parametrize("name,tag", [("medium_data", "medium_data_k6"), ("big_data", "big_data_k12")])
def test_lut_tables(indexes, name, tag):
    _, idx = indexes[name]
    g = gu.load_lut(tag)
    K = int(g["K"])
    table = rp.RefLUT(idx, K).materialise()
    keys = sorted(table)
    assert np.array_equal(np.asarray(keys, np.uint32), g["keys"])
    assert np.array_equal(np.asarray([table[k][0] for k in keys], np.uint32), g["lo"])
    assert np.array_equal(np.asarray([table[k][1] for k in keys], np.uint32), g["hi"])
    if "pos" in g:
        pos = [p for k in keys for p in idx.get_positions(*table[k])]
        assert np.array_equal(np.asarray(pos, np.uint32), g["pos"])
    else:
        pos = [p for k in keys for p in idx.get_positions(*table[k])]
        assert _sha(np.asarray(pos, np.uint32)) == str(g["pos_sha256"])


@pytest.mark.parametrize("name,tag", [("medium_data", "medium_data_k6"), ("big_data", "big_data_k12"), ("big_data", "big_data_k15")])
def test_rmi_lookups(indexes, name, tag):
    _, idx = indexes[name]
    p = gu.load_rmi(tag)
    rmi = rp.RefRMI(idx, p["K"], p["level_sizes"], p["coef"], p["intercept"])
    n_err = 0
    for q, pred, lo, hi in gu.load_json(f"rmi_lookups_{tag}.json.gz"):
        if pred is None:
            with pytest.raises((RecursionError, IndexError)):
                rmi.get_suffix_rmi(q)
            n_err += 1
            continue
        assert rmi.rmi_predict(q) == pred          # bit-identical float64
        assert rmi.get_suffix_rmi(q) == (lo, hi)
    assert n_err <= 5


def _check_set(indexes, fname, stride=1):
    g = gu.load_json(fname)
    _, idx = indexes[g["ref"]]
    s = rp.RefSMEM(idx, lut=rp.RefLUT(idx, g["K_lut"]))
    reads = g["reads"]
    for ml, exp in g["bwa"].items():
        for q, e in list(zip(reads, exp))[::stride]:
            assert gu.norm(s.get_SMEMS(q, int(ml))) == e
    for q, e in list(zip(reads, g["lut"]))[::stride]:
        if e is not None:
            assert gu.norm(s.get_smems_lut(q)) == e
    for tag, exp in g["rmi"].items():
        p = gu.load_rmi(tag)
        s.rmi_lut = rp.RefRMI(idx, p["K"], p["level_sizes"], p["coef"], p["intercept"])
        for q, e in list(zip(reads, exp))[::stride]:
            if e is None:
                continue
            if isinstance(e, dict):
                with pytest.raises((RecursionError, IndexError, TypeError)):
                    s.get_smems_rmi(q)
            else:
                assert gu.norm(s.get_smems_rmi(q)) == e


def test_smems_c1(indexes):
    _check_set(indexes, "smems_c1_big_exact101.json.gz", stride=10)


def test_smems_c2(indexes):
    _check_set(indexes, "smems_c2_big_mixed101.json.gz", stride=13)


def test_smems_sub151(indexes):
    _check_set(indexes, "smems_big_sub151.json.gz", stride=6)


def test_smems_medium_fuzz(indexes):
    _check_set(indexes, "smems_medium_fuzz.json.gz", stride=5)


def test_python_baseline_tool_reports_every_method():
    """bench.py's cpu_baseline.python_port: one process and a pool, three methods, core count stated."""
    import bench
    out = bench.python_port_baseline(0.2)
    assert "unavailable" not in out, out
    assert out["cores"] >= 1
    for m in ("bwa", "lut", "rmi"):
        assert out[f"{m}_reads_per_s_one_process"] > 0 and out[f"{m}_reads_per_s_pool"] > 0
