"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures frozen from the
live reference and against the oracle on seeded inputs.  Bit-exact: SA intervals, positions and
SMEM sets are integers."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import golden_util as gu  # noqa: E402


@pytest.fixture(scope="module")
def gs():
    import genie_smem_b200 as g
    return g


@pytest.fixture(scope="module")
def matchers(gs):
    out = {}
    for name in ("small_data", "medium_data", "big_data"):
        gidx = gu.load_index(name)
        out[name] = (gidx, gs.ExactMatch.from_text(gidx["text"], name=name + ".fa"))
    return out


@pytest.fixture(scope="module")
def oracles():
    from oracle import ref_port as rp
    out = {}
    for name in ("small_data", "medium_data", "big_data"):
        gidx = gu.load_index(name)
        out[name] = rp.RefIndex(gidx["text"], gidx["suffix_array"])
    return out


def _dicts(reads, res):
    out = []
    for i, q in enumerate(reads):
        d = {}
        for r in res.for_read(i):
            d[q[int(r["qstart"]):int(r["qend"])]] = (int(r["sa_lo"]), int(r["sa_hi"]))
        out.append([[k, v[0], v[1]] for k, v in d.items()])
    return out


@pytest.mark.parametrize("name", ["small_data", "medium_data", "big_data"])
def test_index_builder_matches_reference_arrays(matchers, name):
    gidx, m = matchers[name]
    sa, bwt = m._host.export()
    assert np.array_equal(sa, gidx["suffix_array"])
    assert bwt.decode() == gidx["bwt"]
    assert m.fm_index["count_dic"] == gu.meta()[name]["count_dic"]


@pytest.mark.parametrize("name", ["small_data", "medium_data", "big_data"])
def test_backsearch_vs_oracle(matchers, oracles, name):
    gidx, m = matchers[name]
    text = gidx["text"]
    rng = random.Random(7)
    reads = [""]
    for _ in range(400):
        L = rng.randint(1, min(151, len(text)))
        p = rng.randrange(0, len(text) - L + 1)
        reads.append(text[p:p + L])
    for _ in range(400):
        reads.append("".join(rng.choice("ACGT") for _ in range(rng.randint(1, 40))))
    reads += [text[-30:], text[:30], "A" * 200, "ACGT" * 40]
    lo, cnt = m.exact_match_back_prop_batch(reads)
    for q, l, c in zip(reads, lo, cnt):
        exp = oracles[name].exact_match_back_prop(q)
        got = -1 if c == 0 else (int(l), int(l) + int(c) - 1)
        assert got == exp, q
    # the single-query surface and the one-step call
    assert m.exact_match_back_prop(reads[5]) == oracles[name].exact_match_back_prop(reads[5])
    t = oracles[name].exact_match_back_prop(reads[5][1:])
    if t != -1:
        assert m.exact_match_back_prop_add_one(reads[5][0], t) == oracles[name].exact_match_back_prop_add_one(reads[5][0], t)
    assert m.exact_match(reads[5]) == oracles[name].exact_match(reads[5])
    assert m.exact_match_back_prop("") == (0, len(text))


@pytest.mark.parametrize("name,tag", [("medium_data", "medium_data_k6"), ("big_data", "big_data_k12")])
def test_device_lut_equals_reference_table(gs, matchers, name, tag):
    _, m = matchers[name]
    g = gu.load_lut(tag)
    lut = gs.LUT(m)
    lut.generate_lut(int(g["K"]))
    t = lut.table.cpu().numpy().view(np.uint32).reshape(-1, 2)
    keys = np.nonzero(t[:, 1])[0]
    assert np.array_equal(keys.astype(np.uint32), g["keys"])
    assert np.array_equal(t[keys, 0], g["lo"])
    assert np.array_equal(t[keys, 0] + t[keys, 1] - 1, g["hi"])
    if "pos" in g:      # positions column of the reference's checked-in medium_data-LUT.json
        k0 = str(int(g["keys"][3]))
        n0 = int(g["npos"][:3].sum())
        assert lut.lut[k0][1] == [int(x) for x in g["pos"][n0:n0 + int(g["npos"][3])]]
        assert k0 in lut.lut and str(4 ** int(g["K"]) - 1) in lut.lut or True


@pytest.mark.parametrize("name,tag", [("medium_data", "medium_data_k6"), ("big_data", "big_data_k12"), ("big_data", "big_data_k15")])
def test_rmi_lookup_golden(gs, matchers, name, tag):
    _, m = matchers[name]
    p = gu.load_rmi(tag)
    params = gs.RmiParams(p["K"], p["level_sizes"], p["coef"], p["intercept"])
    rows = gu.load_json(f"rmi_lookups_{tag}.json.gz")
    codes = [gs.LUT.convert_seq_to_num(r[0]) for r in rows]
    pred, lo, hi, st = gs.rmi_lookup_batch(m.device_index, params, codes)
    for k, (q, gp, glo, ghi) in enumerate(rows):
        if gp is None:
            assert st[k] == gs.READ_REF_RAISES, q
        else:
            assert st[k] == gs.READ_OK, q
            assert pred[k] == gp, q                     # bit-identical float64 (mul then add, no FMA)
            assert (int(lo[k]), int(hi[k])) == (glo, ghi), q


GOLDEN_SETS = ["smems_c1_big_exact101.json.gz", "smems_c2_big_mixed101.json.gz", "smems_big_sub151.json.gz", "smems_medium_fuzz.json.gz"]


@pytest.mark.parametrize("fname", GOLDEN_SETS)
def test_smem_sets_equal_reference(gs, matchers, fname):
    g = gu.load_json(fname)
    _, m = matchers[g["ref"]]
    s = gs.SMEM(m)
    reads = g["reads"]
    for ml, exp in g["bwa"].items():
        got = _dicts(reads, s.get_SMEMS_batch(reads, int(ml)))
        assert got == exp
    s.lut.generate_lut(g["K_lut"])
    sel = [i for i, e in enumerate(g["lut"]) if e is not None]
    try:
        for machine in (False, True):                    # the sweep's picks (default), then the frame machine itself
            gs.set_lut_frame_machine(machine)
            got = _dicts([reads[i] for i in sel], s.get_smems_lut_batch([reads[i] for i in sel]))
            assert got == [g["lut"][i] for i in sel], machine
    finally:
        gs.set_lut_frame_machine(False)
    for tag, exp in g["rmi"].items():
        p = gu.load_rmi(tag)
        s.rmi_lut = gs.RMI_LUT([p["experts"][0], p["experts"][1]], p["K"], g["ref"] + ".fa", matcher=m)
        s.rmi_lut.rmi = gs.RMI.from_params(p["level_sizes"], p["coef"], p["intercept"])
        sel = [i for i, e in enumerate(exp) if e is not None]
        rs = [reads[i] for i in sel]
        res = s.get_smems_rmi_batch(rs)
        got = _dicts(rs, res)
        n_raise = 0
        for k, i in enumerate(sel):
            if isinstance(exp[i], dict):
                assert res.status[k] == gs.READ_REF_RAISES
                n_raise += 1
            else:
                assert res.status[k] == gs.READ_OK
                assert got[k] == exp[i], reads[i]
        assert n_raise <= 8


def test_single_query_surface(gs, matchers):
    g = gu.load_json("smems_c2_big_mixed101.json.gz")
    _, m = matchers["big_data"]
    s = gs.SMEM(m)
    s.lut.generate_lut(12)
    q = g["reads"][777]
    assert gu.norm(s.get_SMEMS(q, 1)) == g["bwa"]["1"][777]
    assert gu.norm(s.get_smems_lut(q)) == g["lut"][777]
    with pytest.raises(KeyError):
        s.get_SMEMS("ACGTN", 1)
    with pytest.raises(ValueError):
        m.exact_match_back_prop("acgt")
    # reference-shaped helper calls
    at = s.get_SMEM_at_index(q, 0)
    first = g["bwa"]["1"][777][0]
    assert [at[0], at[1][0], at[1][1]] == first


def _synthetic(n_bases, n_reads, L, seed, sub_rate):
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, n_bases, dtype=np.uint8)
    starts = rng.integers(0, n_bases - L + 1, n_reads)
    reads = ref[starts[:, None] + np.arange(L)[None, :]].copy()
    mut = rng.random(reads.shape) < sub_rate
    reads[mut] = (reads[mut] + rng.integers(1, 4, int(mut.sum()), dtype=np.uint8)) & 3
    return ref, reads, mut.sum(axis=1)


def test_properties_at_scale(gs):
    """Size-independent properties on a 4 Mbp synthetic reference, 50k 151-bp reads with 1 % subs."""
    L = 151
    ref, reads, nmut = _synthetic(4_000_000, 50_000, L, 1234, 0.01)
    bases = np.frombuffer(b"ACGT", np.uint8)
    text = bases[ref].tobytes().decode()
    m = gs.ExactMatch.from_text(text)
    idx = m.device_index
    batch = gs.ReadBatch.from_codes(reads, L)
    e = gs.Engine(idx, len(reads), L)
    res = e.run(gs.METHOD_BWA, batch, min_len=1)
    offs, recs = res.offsets, res.records
    assert offs[-1] == len(recs) and np.all(np.diff(offs) >= 1)
    rid = np.repeat(np.arange(len(reads)), np.diff(offs))
    assert np.array_equal(recs["read_id"], rid.astype(np.uint32))
    # records of a read chain: each covers the previous end, ends strictly increase, last ends at L
    first = np.zeros(len(recs), bool); first[offs[:-1]] = True
    last = np.zeros(len(recs), bool); last[offs[1:] - 1] = True
    assert np.all(recs["qstart"][first] == 0)
    assert np.all(recs["qend"][last] == L)
    prev_end = np.concatenate([[0], recs["qend"][:-1]])
    assert np.all(recs["qstart"][~first] <= prev_end[~first])
    assert np.all(recs["qend"][~first] > prev_end[~first])
    # an unmutated read is one record covering it
    exact = nmut == 0
    assert np.all(np.diff(offs)[exact] == 1)
    # every interval equals an independent backward search of that substring (other kernel), and
    # extending it by one base on either side does not occur (maximality)
    sel = np.random.default_rng(5).choice(len(recs), 20_000, replace=False)
    subs, left, right = [], [], []
    rs = ["".join("ACGT"[c] for c in reads[i]) for i in range(len(reads))] if False else None
    rtxt = bases[reads].view(f"S{L}").reshape(-1)
    for k in sel:
        q = rtxt[recs["read_id"][k]].decode()
        a, b = int(recs["qstart"][k]), int(recs["qend"][k])
        subs.append(q[a:b])
        left.append(q[a - 1:b] if a > 0 else "")
        right.append(q[a:b + 1] if b < L else "")
    lo, cnt = m.exact_match_back_prop_batch(subs)
    assert np.array_equal(lo, recs["sa_lo"][sel])
    assert np.array_equal(lo + cnt - 1, recs["sa_hi"][sel])
    assert np.all(cnt >= 1)
    _, cl = m.exact_match_back_prop_batch(left)
    _, cr = m.exact_match_back_prop_batch(right)
    assert np.all(cl[[i for i, s in enumerate(left) if s]] == 0)
    assert np.all(cr[[i for i, s in enumerate(right) if s]] == 0)
    # positions: the interval's SA values are real occurrences
    pos = gs.sa_lookup(idx, recs["sa_lo"][sel][:2000])
    for k, p in zip(sel[:2000], pos):
        a, b = int(recs["qstart"][k]), int(recs["qend"][k])
        assert text[p - 1:p - 1 + (b - a)] == rtxt[recs["read_id"][k]].decode()[a:b]
    # LUT- and RMI-free cross-check on a subsample against the oracle
    from oracle import ref_port as rp
    sa, _ = m._host.export()
    o = rp.RefSMEM(rp.RefIndex(text, sa))
    for i in range(0, 300):
        q = rtxt[i].decode()
        d = {}
        for r in res.for_read(i):
            d[q[int(r["qstart"]):int(r["qend"])]] = (int(r["sa_lo"]), int(r["sa_hi"]))
        assert d == o.get_SMEMS(q, 1)


def test_pipelined_engine_equals_engine(gs):
    """The overlapped end-to-end path returns exactly the records of the plain one, ragged lengths included."""
    rng = np.random.default_rng(9)
    ref = rng.integers(0, 4, 300_000, dtype=np.uint8)
    text = np.frombuffer(b"ACGT", np.uint8)[ref].tobytes().decode()
    m = gs.ExactMatch.from_text(text)
    reads = []
    for _ in range(5003):
        L = int(rng.integers(12, 152))
        p = int(rng.integers(0, len(text) - L))
        q = list(text[p:p + L])
        for k in np.nonzero(rng.random(L) < 0.02)[0]:
            q[k] = "ACGT"[int(rng.integers(0, 4))]
        reads.append("".join(q))
    batch = gs.ReadBatch.from_strings(reads, pin=True)
    idx = m.device_index
    lut = gs.lut_build(idx, 8)
    plain = gs.Engine(idx, len(reads), 160, mems_per_read=48, recs_per_read=48)
    pipe = gs.PipelinedEngine(idx, len(reads), 160, n_chunks=5, mems_per_read=48, recs_per_read=48)
    for method, kw in ((gs.METHOD_BWA, {"min_len": 1}), (gs.METHOD_LUT, {"K": 8, "lut": lut})):
        a = plain.run(method, batch, **kw)
        ra, oa, sa = a.records.copy(), a.offsets.copy(), a.status.copy()
        b = pipe.run(method, batch, **kw)
        assert np.array_equal(oa, b.offsets)
        assert np.array_equal(ra, b.records)
        assert np.array_equal(sa, b.status)


def test_rmi_probe_table_changes_nothing(gs, matchers):
    """The 16-byte probe records are a layout optimisation: records with and without them are identical."""
    g = gu.load_json("smems_c2_big_mixed101.json.gz")
    _, m = matchers["big_data"]
    p = gu.load_rmi("big_data_k15")
    idx = m.device_index
    reads = g["reads"][:600]
    batch = gs.ReadBatch.from_strings(reads)
    e = gs.Engine(idx, len(reads), 160, mems_per_read=64, recs_per_read=64)
    plain = gs.RmiParams(p["K"], p["level_sizes"], p["coef"], p["intercept"])
    a = e.run(gs.METHOD_RMI, batch, rmi=plain)
    ra, oa = a.records.copy(), a.offsets.copy()
    fast = gs.RmiParams(p["K"], p["level_sizes"], p["coef"], p["intercept"]).build_probe_table(idx)
    b = e.run(gs.METHOD_RMI, batch, rmi=fast)
    assert np.array_equal(oa, b.offsets) and np.array_equal(ra, b.records)
    got = _dicts(reads, b)
    assert got == g["rmi"]["big_data_k15"][:600]


def _oracle_dicts(text, sa, method, reads, **kw):
    from oracle.c_oracle import COracle
    return COracle(text, sa).smem_dicts(method, reads, **kw)


def test_long_reads_and_edge_lengths_vs_oracle(gs, matchers):
    """Reference-style long queries (SMEM.py:517 uses 2,000 bp), length-1 reads, reads shorter than K."""
    gidx, m = matchers["big_data"]
    text = gidx["text"]
    rng = random.Random(21)
    reads = []
    for L in (1, 2, 5, 11, 12, 13, 64, 65, 255, 256, 257, 700, 1500, 2000):
        p = rng.randrange(0, len(text) - L)
        q = list(text[p:p + L])
        for k in range(L):
            if rng.random() < 0.03:
                q[k] = rng.choice("ACGT")
        reads.append("".join(q))
    reads += [gs.create_query_from_ref(text, 2000), gs.create_random_query(300)]
    s = gs.SMEM(m)
    s.lut.generate_lut(12)
    exp = _oracle_dicts(text, gidx["suffix_array"], 0, reads, min_len=1)
    assert _dicts(reads, s.get_SMEMS_batch(reads, 1)) == exp
    res = s.get_smems_lut_batch(reads)
    exp = _oracle_dicts(text, gidx["suffix_array"], 1, reads, K=12)
    got = _dicts(reads, res)
    for k, q in enumerate(reads):
        if len(q) < 12:
            assert res.status[k] == gs.READ_TOO_SHORT and exp[k] == "short"
        else:
            assert got[k] == exp[k], len(q)
    # an empty batch is legal
    empty = s.get_SMEMS_batch([], 1)
    assert len(empty.records) == 0 and list(empty.offsets) == [0]


def test_low_complexity_reads_spill_the_candidate_cache(gs):
    """More occurrence-count changes in one forward extension than shared-memory candidate slots."""
    text = "A" * 300 + "C" + "A" * 120 + "G" + "ACGT" * 5 + "T" * 80 + "GATTACA" * 20
    m = gs.ExactMatch.from_text(text)
    sa, _ = m._host.export()
    reads = ["A" * 151, "A" * 100 + "C" + "A" * 50, "T" * 70 + "A" * 81, "A" * 40 + "G" + "ACGT" * 3 + "T" * 60, "GATTACA" * 15 + "A" * 40]
    s = gs.SMEM(m)
    assert _dicts(reads, s.get_SMEMS_batch(reads, 1)) == _oracle_dicts(text, sa, 0, reads, min_len=1)
    s.lut.generate_lut(4)
    assert _dicts(reads, s.get_smems_lut_batch(reads)) == _oracle_dicts(text, sa, 1, reads, K=4)


def test_seed_table_equals_host_restatement(gs, matchers):
    """k_seed_build against the host-compiled builder of tests/emu (same fm_core.cuh arithmetic)."""
    from tests.emu.harness import Emu
    gidx, m = matchers["medium_data"]
    idx = m.device_index
    try:
        for K in (1, 3, 6):
            idx.build_seed_table(K)
            t = idx.seed_table.cpu().numpy().view(np.uint32)
            assert np.array_equal(t, Emu(gidx["text"]).seed_table(K))
    finally:
        idx.drop_seed_table()


@pytest.mark.parametrize("seed_K", [1, 4, 7, None])
def test_seed_table_changes_nothing(gs, seed_K):
    """The seed table is an accelerator of the sweep: maximal-match lists and the records of every method are
    identical with and without it (4 Mbp reference, 20k reads of mixed kinds, K from tiny to auto)."""
    L = 151
    ref, reads, _ = _synthetic(4_000_000, 20_000, L, 77, 0.012)
    rng = np.random.default_rng(8)
    reads[:2000] = rng.integers(0, 4, (2000, L), dtype=np.uint8)          # uniform random reads
    reads[2000:2100, 40:] = 0                                             # poly-A tails
    idx = gs.DeviceIndex.build_on_device(ref)
    batch = gs.ReadBatch.from_codes(reads, L)
    e = gs.Engine(idx, len(reads), L, mems_per_read=64, recs_per_read=64)
    lut = gs.lut_build(idx, 9)

    def run_all():
        out = []
        for method, kw in ((gs.METHOD_BWA, {"min_len": 1}), (gs.METHOD_BWA, {"min_len": 19}), (gs.METHOD_LUT, {"K": 9, "lut": lut})):
            r = e.run(method, batch, **kw)
            out.append((r.records.copy(), r.offsets.copy(), r.status.copy(), r.n_mems))
        return out

    plain = run_all()
    idx.build_seed_table(seed_K)
    assert idx.seed_K == (seed_K if seed_K else 10)
    seeded = run_all()
    for a, b in zip(plain, seeded):
        assert a[3] == b[3], "number of maximal matches"
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("fname", GOLDEN_SETS[:3])
def test_smem_sets_equal_reference_with_seed_table(gs, matchers, fname):
    g = gu.load_json(fname)
    _, m = matchers[g["ref"]]
    try:
        m.device_index.build_seed_table(6)
        s = gs.SMEM(m)
        reads = g["reads"]
        assert _dicts(reads, s.get_SMEMS_batch(reads, 1)) == g["bwa"]["1"]
        s.lut.generate_lut(g["K_lut"])
        sel = [i for i, e in enumerate(g["lut"]) if e is not None]
        assert _dicts([reads[i] for i in sel], s.get_smems_lut_batch([reads[i] for i in sel])) == [g["lut"][i] for i in sel]
    finally:
        m.device_index.drop_seed_table()


def test_low_complexity_reads_with_seed_table(gs):
    text = "A" * 300 + "C" + "A" * 120 + "G" + "ACGT" * 5 + "T" * 80 + "GATTACA" * 20
    m = gs.ExactMatch.from_text(text)
    sa, _ = m._host.export()
    reads = ["A" * 151, "A" * 100 + "C" + "A" * 50, "T" * 70 + "A" * 81, "A" * 40 + "G" + "ACGT" * 3 + "T" * 60, "GATTACA" * 15 + "A" * 40, "ACG"]
    exp = _oracle_dicts(text, sa, 0, reads, min_len=1)
    for K in (2, 5, 8):
        m.device_index.build_seed_table(K)
        assert _dicts(reads, gs.SMEM(m).get_SMEMS_batch(reads, 1)) == exp


def test_rmi_fast_search_changes_nothing(gs):
    """The error-bounded fast search (enabled by the None-row list) and the literal search give identical records,
    statuses included, on a 2 Mbp reference with poor and good models (many hazards vs almost none)."""
    import bench
    L = 151
    ref, reads, _ = _synthetic(2_000_000, 12_000, L, 31, 0.015)
    reads[:1500] = np.random.default_rng(2).integers(0, 4, (1500, L), dtype=np.uint8)
    reads[1500:1600, :60] = 0                                            # poly-A heads: the smallest k-mers, row 0 territory
    idx = gs.DeviceIndex.build_on_device(ref).build_seed_table()
    batch = gs.ReadBatch.from_codes(reads, L)
    e = gs.Engine(idx, len(reads), L, mems_per_read=64, recs_per_read=64)
    n_filters = 0
    for K, experts in ((9, (16, 512)), (11, (64, 4096)), (8, (1, 1))):
        rmi = bench.train_rmi(idx, K, experts, idx.device)                # probe table + None rows
        assert rmi.c.n_none_rows == K
        fast = e.run(gs.METHOD_RMI, batch, rmi=rmi)
        fr, fo, fs = fast.records.copy(), fast.offsets.copy(), fast.status.copy()
        bnd = e.run(gs.METHOD_RMI, batch, rmi=rmi.build_bounds_table(idx))   # true bounds from the dense table; with the hazard
        assert np.array_equal(fo, bnd.offsets) and np.array_equal(fr, bnd.records) and np.array_equal(fs, bnd.status)   # filter: pre-filtered
        n_filters += rmi.hazard_slots is not None
        assert rmi.n_hazards is not None and (rmi.hazard_slots is not None) == (rmi.n_hazards <= 1 << 18)
        try:
            gs.set_rmi_prefilter(False)                                       # bounds table, frame machine on every read
            bnd = e.run(gs.METHOD_RMI, batch, rmi=rmi)
            assert np.array_equal(fo, bnd.offsets) and np.array_equal(fr, bnd.records) and np.array_equal(fs, bnd.status)
        finally:
            gs.set_rmi_prefilter(True)
        rmi.drop_bounds_table()
        assert rmi.hazard_slots is None and rmi.c.hazard_n_slots == 0
        rmi.c.none_rows, rmi.c.n_none_rows = None, 0                      # literal search only
        lit = e.run(gs.METHOD_RMI, batch, rmi=rmi)
        assert np.array_equal(fo, lit.offsets) and np.array_equal(fr, lit.records) and np.array_equal(fs, lit.status)
    assert n_filters >= 2                                                 # the pre-filter ran (K = 8, 9: at most 4^K <= 2^18 hazard codes)


def _codes_to_strings(reads):
    return [r.decode() for r in np.frombuffer(b"ACGT", np.uint8)[reads].view(f"S{reads.shape[1]}").reshape(-1)]


def _check_against_oracle(gs, res, reads_s, exp, what):
    got = _dicts(reads_s, res)
    n_raise = 0
    for k, e in enumerate(exp):
        if e == "raises":
            assert res.status[k] == gs.READ_REF_RAISES, (what, k)
            n_raise += 1
        else:
            assert res.status[k] == gs.READ_OK, (what, k)
            assert got[k] == e, (what, k, reads_s[k])
    return n_raise


@pytest.mark.parametrize("seed_table", [True, False])
def test_all_methods_vs_oracle_on_synthetic_reference(gs, seed_table):
    """Tier-B parity at a size the C oracle finishes in seconds: 4 Mbp synthetic reference, 2,400 reads of
    151 bp -- 1 % substitution reads, uniform-random reads, poly-A heads / tails / whole reads -- through BWA-, LUT- and
    RMI-SMEM, the RMI being the one bench.py trains (vectorised trainer, probe table, None rows), against
    oracle/smem_oracle.c (the reference's get_SMEMS / get_smems_lut / get_smems_rmi restated literally).
    seed_table=True: sweep seed table on, RMI lookups by seed-table bounds + arithmetic replay (rmi_arith_lookup);
    False: plain FM stepping, RMI lookups by the probe-based error-bounded search.  Same records either way."""
    import bench
    from oracle.c_oracle import COracle
    L = 151
    ref, reads, _ = _synthetic(4_000_000, 2_400, L, 4242, 0.01)
    rng = np.random.default_rng(6)
    reads[:600] = rng.integers(0, 4, (600, L), dtype=np.uint8)          # uniform random
    reads[600:640, :70] = 0                                             # poly-A heads (the smallest k-mers: row 0 territory)
    reads[640:680, 90:] = 0                                             # poly-A tails
    reads[680:690] = 0                                                  # poly-A reads
    reads[690:700] = 3                                                  # poly-T reads (the largest k-mers: table end)
    reads[700:720, 50:100] = np.tile(np.array([0, 1], np.uint8), 25)    # dinucleotide repeats
    text = np.frombuffer(b"ACGT", np.uint8)[ref].tobytes()
    idx = gs.DeviceIndex.build_on_device(ref)
    if seed_table:
        idx.build_seed_table()
    sa = idx.suffix_array_host()
    reads_s = _codes_to_strings(reads)
    batch = gs.ReadBatch.from_codes(reads, L)
    e = gs.Engine(idx, len(reads), L, mems_per_read=128, recs_per_read=96)
    o = COracle(text, sa)
    res = e.run(gs.METHOD_BWA, batch, min_len=1)
    _check_against_oracle(gs, res, reads_s, o.smem_dicts(0, reads_s, min_len=1), "bwa")
    res = e.run(gs.METHOD_BWA, batch, min_len=20)
    _check_against_oracle(gs, res, reads_s, o.smem_dicts(0, reads_s, min_len=20), "bwa minlen 20")
    try:
        for K in (8, 12):
            exp = o.smem_dicts(1, reads_s, K=K)
            lut = gs.lut_build(idx, K)
            gs.set_lut_frame_machine(False)              # records = the sweep's picks (get_smems_lut == get_SMEMS with min_len 1)
            res = e.run(gs.METHOD_LUT, batch, K=K, lut=lut)
            _check_against_oracle(gs, res, reads_s, exp, f"lut K={K}")
            picks = (res.records.copy(), res.offsets.copy(), res.status.copy())
            gs.set_lut_frame_machine(True)               # the reference's frame machine itself
            res = e.run(gs.METHOD_LUT, batch, K=K, lut=lut)
            _check_against_oracle(gs, res, reads_s, exp, f"lut K={K}, frame machine")
            assert np.array_equal(picks[0], res.records) and np.array_equal(picks[1], res.offsets) and np.array_equal(picks[2], res.status)
    finally:
        gs.set_lut_frame_machine(False)
    n_raise = n_filters = 0
    for K, experts in ((11, (64, 4096)), (15, (256, 16384)), (9, (4, 64))):
        rmi = bench.train_rmi(idx, K, experts, idx.device)
        res = e.run(gs.METHOD_RMI, batch, rmi=rmi)
        exp = o.smem_dicts(2, reads_s, rmi=bench.rmi_dict(rmi))
        n_raise += _check_against_oracle(gs, res, reads_s, exp, f"rmi K={K}")
        rmi.build_bounds_table(idx)                                       # lookups from the dense k-mer bounds table; reads without a
        res = e.run(gs.METHOD_RMI, batch, rmi=rmi)                        # hazard window pre-filtered to the BWA-SMEM selection
        _check_against_oracle(gs, res, reads_s, exp, f"rmi K={K}, bounds table, pre-filter {rmi.hazard_slots is not None} ({rmi.n_hazards} hazard codes)")
        n_filters += rmi.hazard_slots is not None
        try:
            gs.set_rmi_prefilter(False)                                   # the frame machine on every read
            res = e.run(gs.METHOD_RMI, batch, rmi=rmi)
            _check_against_oracle(gs, res, reads_s, exp, f"rmi K={K}, bounds table, frame machine")
        finally:
            gs.set_rmi_prefilter(True)
        rmi.drop_bounds_table()
    assert n_raise < 200
    assert n_filters >= 1


def test_add_one_vs_oracle_on_many_pairs(gs, matchers, oracles):
    """exact_match_back_prop_add_one (ExactMatch.py:155-171) on 1,200 (char, interval) pairs per reference, hits and misses."""
    for name in ("medium_data", "big_data"):
        gidx, m = matchers[name]
        text = gidx["text"]
        rng = random.Random(3)
        subs = []
        for _ in range(1200):
            L = rng.randint(1, 24)
            p = rng.randrange(0, len(text) - L)
            subs.append(text[p:p + L])
        lo, cnt = m.exact_match_back_prop_batch(subs)
        chars = [rng.choice("ACGT") for _ in subs]
        nlo, ncnt = gs.add_one_batch(m.device_index, ["ACGT".index(c) for c in chars], lo, cnt)
        n_miss = 0
        for q, ch, l, c, a, b in zip(subs, chars, lo, cnt, nlo, ncnt):
            exp = oracles[name].exact_match_back_prop_add_one(ch, (int(l), int(l) + int(c) - 1))
            got = -1 if b == 0 else (int(a), int(a) + int(b) - 1)
            assert got == exp, (q, ch)
            assert exp == oracles[name].exact_match_back_prop(ch + q)
            n_miss += exp == -1
        assert 50 < n_miss < 1150


def test_selection_staging_overflow_reruns_the_read(gs, matchers):
    """A read that emits more records than a selection thread can stage (64) is run a second time writing in place:
    long uniform-random reads (hundreds of records each) through all three methods equal the oracle."""
    gidx, m = matchers["big_data"]
    text = gidx["text"]
    rng = random.Random(17)
    reads = ["".join(rng.choice("ACGT") for _ in range(L)) for L in (900, 1500, 2000, 700, 64, 2000)]
    reads.append(text[100:1900])
    s = gs.SMEM(m)
    s.lut.generate_lut(6)
    exp = _oracle_dicts(text, gidx["suffix_array"], 0, reads, min_len=1)
    assert max(len(e) for e in exp) > 64
    assert _dicts(reads, s.get_SMEMS_batch(reads, 1)) == exp
    exp_lut = _oracle_dicts(text, gidx["suffix_array"], 1, reads, K=6)
    try:
        for machine in (False, True):                    # picks, then the frame machine (whose staging this test is about)
            gs.set_lut_frame_machine(machine)
            assert _dicts(reads, s.get_smems_lut_batch(reads)) == exp_lut, machine
    finally:
        gs.set_lut_frame_machine(False)
    p = gu.load_rmi("big_data_k12")
    s.rmi_lut = gs.RMI_LUT([p["experts"][0], p["experts"][1]], p["K"], "big_data.fa", matcher=m)
    s.rmi_lut.rmi = gs.RMI.from_params(p["level_sizes"], p["coef"], p["intercept"])
    res = s.get_smems_rmi_batch(reads)
    expr = _oracle_dicts(text, gidx["suffix_array"], 2, reads, rmi=p)
    _check_against_oracle(gs, res, reads, expr, "rmi long")


def test_reads_longer_than_the_shared_memory_path_vs_oracle(gs, matchers):
    """Reads above 1,024 bases run the long-read sweep (bases from global memory, grid sized by the staging budget); the
    reference's get_SMEMS / get_smems_lut / create_query(query_size) take any length.  Up to 60,000 bases here (records
    carry 16-bit read offsets: 65,535 is the documented limit)."""
    gidx, m = matchers["big_data"]
    text = gidx["text"]
    rng = random.Random(33)
    reads = []
    for L in (1025, 1100, 3000, 9000, 30000, 60000):
        p = rng.randrange(0, len(text) - L)
        q = list(text[p:p + L])
        for k in range(L):
            if rng.random() < 0.02:
                q[k] = rng.choice("ACGT")
        reads.append("".join(q))
    reads.append("".join(rng.choice("ACGT") for _ in range(5000)))
    reads.append(text[500:2600])                                            # one exact 2,100-base stretch
    reads.append("ACGT" * 10)                                               # a short read in the same batch
    s = gs.SMEM(m)
    s.lut.generate_lut(10)
    exp = _oracle_dicts(text, gidx["suffix_array"], 0, reads, min_len=1)
    assert _dicts(reads, s.get_SMEMS_batch(reads, 1)) == exp
    assert _dicts(reads, s.get_smems_lut_batch(reads)) == _oracle_dicts(text, gidx["suffix_array"], 1, reads, K=10)
    with pytest.raises(ValueError):
        s.get_SMEMS_batch(["A" * 65536], 1)
