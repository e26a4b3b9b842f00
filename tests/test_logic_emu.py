"""Host build of the kernels' per-read logic (tests/emu: the SAME sweep_logic.cuh / select_logic.cuh
/ fm_core.cuh the CUDA kernels compile) against the golden fixtures and the oracle.  This is what
keeps the control flow honest in the GPU-less container; the GPU runs the same headers in
tests/test_gpu_parity.py."""
import random

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import ref_port as rp
from tests import golden_util as gu
from tests.emu.harness import Emu, records_to_dict


@pytest.fixture(scope="module")
def emus():
    out = {}
    for name in ("small_data", "medium_data", "big_data"):
        g = gu.load_index(name)
        out[name] = (g, Emu(g["text"]))
    return out


def test_device_layout_lut_equals_reference_table(emus):
    _, em = emus["medium_data"]
    g = gu.load_lut("medium_data_k6")
    t = em.lut(6).reshape(-1, 2)
    keys = np.nonzero(t[:, 1])[0]
    assert np.array_equal(keys.astype(np.uint32), g["keys"])
    assert np.array_equal(t[keys, 0], g["lo"])
    assert np.array_equal(t[keys, 0] + t[keys, 1] - 1, g["hi"])


@pytest.mark.parametrize("tag,name", [("medium_data_k6", "medium_data"), ("big_data_k15", "big_data")])
def test_rmi_lookup_logic(emus, tag, name):
    _, em = emus[name]
    p = gu.load_rmi(tag)
    n_fast = n_arith = 0
    for q, pred, lo, hi in gu.load_json(f"rmi_lookups_{tag}.json.gz")[::3]:
        code = 0
        for ch in q:
            code = code << 2 | "ACGT".index(ch)
        s, gp, glo, ghi = em.rmi_lookup(p, code)
        s2, lo2, hi2, _ = em.rmi_search(p, code)          # the resumable machine the kernels run
        hz, lo3, hi3, _ = em.rmi_fast_lookup(p, code)     # the error-bounded fast search (defers on a hazard)
        hz4, lo4, hi4 = em.rmi_arith_lookup(p, code)      # the same without probes (bounds from the FM index)
        n_fast += not hz
        n_arith += not hz4
        if pred is None:
            assert s == -1 and s2 == -1 and hz and hz4
        else:
            assert s == 0 and gp == pred and (glo, ghi) == (lo, hi)
            assert s2 == 0 and (lo2, hi2) == (lo, hi)
            assert hz or (lo3, hi3) == (lo, hi), q
            assert hz4 or (lo4, hi4) == (lo, hi), q
    assert n_fast > 0.8 * len(gu.load_json(f"rmi_lookups_{tag}.json.gz")[::3])
    assert n_arith > 0.8 * len(gu.load_json(f"rmi_lookups_{tag}.json.gz")[::3])


def test_lut_smem_records_equal_bwa_smem_records(emus):
    """get_smems_lut emits exactly the records of get_SMEMS(min_len 1) for reads of at least K bases (what lets
    gsm_smem_select(LUT) use the sweep's picks): the frame machine (Selector::run_seeded) against Selector::run_bwa at RECORD
    level -- emission order and duplicates included -- on mutated substrings, random reads and low-complexity reads of the
    three golden references, K = 2..10."""
    n = 0
    for name, (g, em) in emus.items():
        text = g["text"]
        rng = random.Random(7 * len(text))
        reads = []
        for _ in range(150):
            L = rng.choice((12, 25, 60, 101, 151))
            L = min(L, len(text) - 1)
            p = rng.randrange(0, max(1, len(text) - L))
            q = list(text[p:p + L])
            pm = rng.choice((0.0, 0.02, 0.08, 0.3))
            for k in range(len(q)):
                if rng.random() < pm:
                    q[k] = rng.choice("ACGT")
            reads.append("".join(q))
        reads += ["".join(rng.choice("ACGT") for _ in range(rng.choice((10, 40, 101)))) for _ in range(40)]
        reads += ["A" * 50, "A" * 20 + "C" + "A" * 29, "ACGT" * 15, "AC" * 30 + "G" + "AC" * 10, "T" * 12]
        for K in (2, 4, 6, 8, 10):
            for q in reads:
                if len(q) < K:
                    assert em.smem(1, q, K=K) == "short"
                    continue
                assert em.smem(1, q, K=K) == em.smem(0, q, min_len=1), (name, K, q)
                n += 1
    assert n > 2500


def test_get_smems_lut_equals_get_SMEMS_on_the_literal_restatement():
    """The same identity on oracle/ref_port.py (the reference's two routines restated line by line, pinned to the golden
    vectors): get_smems_lut(q) == get_SMEMS(q, 1) as ordered dicts on ~15,000 adversarial small cases -- periodic, skewed and
    palindromic references, reads with ties everywhere, K = 1..8.  (The same generator ran 11.5 M cases without a difference.)"""
    import genie_smem_b200 as gs
    rnd = random.Random(2024)
    n = 0
    for _ in range(130):
        n_ref = rnd.choice((20, 40, 80, 160, 400))
        mode = rnd.random()
        alpha = rnd.choice(("ACGT", "ACGT", "AC", "AAAC", "AAAAAACG", "ACG"))
        if mode < 0.35:
            text = "".join(rnd.choice(alpha) for _ in range(n_ref))
        elif mode < 0.75:
            unit = "".join(rnd.choice("ACGT") for _ in range(rnd.choice((1, 2, 3, 4, 5, 7, 11))))
            t = list((unit * (n_ref // len(unit) + 1))[:n_ref])
            pm = rnd.choice((0.0, 0.02, 0.1))
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
        for K in rnd.sample((1, 2, 3, 4, 5, 6, 8), 3):
            o = rp.RefSMEM(idx, lut=rp.RefLUT(idx, K))
            for _ in range(40):
                L = rnd.choice((K, K + 1, K + 2, 2 * K, 12, 25, 60))
                r = rnd.random()
                if r < 0.6:
                    L = min(L, len(text) - 1)
                    p = rnd.randrange(0, max(1, len(text) - L))
                    q = list(text[p:p + L])
                    pm = rnd.choice((0.0, 0.03, 0.1, 0.3))
                    for k in range(len(q)):
                        if rnd.random() < pm:
                            q[k] = rnd.choice("ACGT")
                    q = "".join(q)
                elif r < 0.8:
                    q = "".join(rnd.choice(alpha) for _ in range(L))
                else:
                    q = "".join(rnd.choice("ACGT") for _ in range(L))
                if len(q) < K or not q:
                    continue
                a = o.get_SMEMS(q, 1)
                b = {k: tuple(v) for k, v in o.get_smems_lut(q).items()}
                assert a == b and list(a) == list(b), (text, q, K)
                n += 1
    assert n > 12000


def test_hand_over_picks_equal_run_bwa(emus):
    """The BWA-SMEM picks the sweep's hand-over makes (one maximum of bwa_pick_key per pick) select exactly the records of
    Selector::run_bwa (get_SMEMS, SMEM.py:456-467) -- on the golden reads of the three references, on low-complexity reads,
    and on random lists with gaps (positions no match covers: outside the reference's domain, still the same picks)."""
    n_checked = 0
    for name, (g, em) in emus.items():
        text = g["text"]
        rng = random.Random(len(text))
        reads = []
        for _ in range(250):
            L = rng.choice((30, 60, 101, 151))
            p = rng.randrange(0, max(1, len(text) - L))
            q = list(text[p:p + L])
            for k in range(len(q)):
                if rng.random() < 0.03:
                    q[k] = rng.choice("ACGT")
            reads.append("".join(q))
        reads += ["A" * 40 + "C" + "A" * 30, "ACGT" * 20, "".join(rng.choice("ACGT") for _ in range(120))]
        for q in reads:
            mems, _ = em.sweep(q)
            if not 1 <= len(mems) <= 32:
                continue
            mask = em.bwa_picks(mems)
            picked = [(m[0], m[1], m[2], m[2] + m[3] - 1) for k, m in enumerate(mems) if (mask >> k) & 1]
            assert picked == em.smem(0, q, min_len=1), (name, q)
            n_checked += 1
    assert n_checked > 600
    _, em = emus["small_data"]
    rng = random.Random(5)
    for _ in range(20000):                                  # lists with gaps, against the literal loop of Selector::run_bwa
        L = rng.randint(1, 60)
        n = rng.randint(1, min(12, L))
        ss, es = sorted(rng.sample(range(0, L), n)), sorted(rng.sample(range(1, L + 1), n))
        mems = [(s, e) for s, e in zip(ss, es) if e > s]
        if not mems:
            continue
        exp, p, frm = 0, 0, 0
        while p < L and frm < len(mems):
            while frm < len(mems) and mems[frm][1] <= p:
                frm += 1
            if frm >= len(mems):
                break
            best, bl = frm, 0
            for k in range(frm, len(mems)):
                if mems[k][0] > p:
                    break
                if mems[k][1] - mems[k][0] > bl:
                    bl, best = mems[k][1] - mems[k][0], k
            exp |= 1 << best
            p = mems[best][1]
        assert em.bwa_picks(mems) == exp, mems


def test_closed_form_gallops_equal_the_probe_loops(emus):
    """rmi_arith_lookup computes the two exponential phases of RMI_LUT.exponential_search (RMI_LUT.py:151-178) in closed
    form; against the probe-by-probe loops on 6 M random and edge-case (start, bounds) triples: tiny and 2^31-row tables,
    None rows at both table ends, starts around both bounds, prediction outside the table."""
    _, em = emus["small_data"]
    for n_rows, none in ((1, [0]), (2, [1]), (7, [0, 3]), (1000, [0, 1, 999]), (100_001, [5, 50_000, 100_000]),
                         (2_147_483_648, [0, 17, 2_147_483_647]), (4_294_967_295, [1, 4_294_967_294])):
        nr = np.asarray(none, np.uint32)
        assert em.lib.emu_rmi_arith_fuzz(n_rows, len(nr), nr.ctypes.data, 1_000_000 if n_rows > 7 else 200_000, n_rows) == 0, n_rows


def test_rmi_search_machine_equals_literal_search_on_bad_models(emus):
    """Differential test of RmiSearch against the literal RmiTable on deliberately poor models: predictions far off,
    negative, beyond the table -- the paths where the reference wraps (negative rows), skips None rows or raises."""
    g, em = emus["medium_data"]
    rng = random.Random(11)
    n = em.n_rows
    for trial in range(60):
        K = rng.choice([3, 6, 9])
        scale = rng.choice([0.0, 0.3, 1.0, 1.7]) * n / 4 ** K
        rmi = {"K": K, "level_sizes": [1, 4, 1][:rng.choice([1, 3])] if False else [1], "coef": [scale], "intercept": [rng.uniform(-1.5 * n, 1.5 * n) if trial % 3 == 0 else rng.uniform(-40, 40)]}
        for _ in range(150):
            code = rng.randrange(4 ** K)
            a = em.rmi_lookup(rmi, code)
            b = em.rmi_search(rmi, code)
            f = em.rmi_fast_lookup(rmi, code)
            h = em.rmi_arith_lookup(rmi, code)
            assert (a[0] == -1) == (b[0] == -1), (rmi, code)
            if a[0] == 0:
                assert (a[2], a[3]) == (b[1], b[2]), (rmi, code)
                assert f[0] or (f[1], f[2]) == (a[2], a[3]), (rmi, code)     # no hazard => identical bounds
                assert h[0] or (h[1], h[2]) == (a[2], a[3]), (rmi, code)
            else:
                assert f[0] and h[0], (rmi, code)                            # the reference raises => never the fast paths


@pytest.mark.parametrize("seed_K", [0, 5, 9])
@pytest.mark.parametrize("fname,stride", [("smems_c1_big_exact101.json.gz", 7), ("smems_c2_big_mixed101.json.gz", 3),
                                          ("smems_big_sub151.json.gz", 2), ("smems_medium_fuzz.json.gz", 2)])
def test_sweep_and_select_logic_vs_reference(emus, fname, stride, seed_K):
    """seed_K: K of the sweep's seed table (0 = plain FM stepping); the records must not depend on it."""
    g = gu.load_json(fname)
    _, em = emus[g["ref"]]
    em.seed_K = seed_K
    em.uniq = {0: 0, 5: 1, 9: 2}[seed_K]  # unique-match shortcut of sweep_logic.cuh: off / forward only (the kernel's) / both directions
    if seed_K:
        stride *= 2
    reads = g["reads"]
    for ml, exp in g["bwa"].items():
        for q, e in list(zip(reads, exp))[::stride]:
            assert records_to_dict(q, em.smem(0, q, int(ml))) == e
    for q, e in list(zip(reads, g["lut"]))[::stride]:
        if e is not None:
            assert records_to_dict(q, em.smem(1, q, 1, g["K_lut"])) == e
    for tag, exp in g["rmi"].items():
        p = gu.load_rmi(tag)
        em.rmi_fast = seed_K != 0          # half of the runs go through the error-bounded fast search first
        for q, e in list(zip(reads, exp))[::stride]:
            if e is None:
                continue
            r = em.smem(2, q, 1, 0, p)
            if isinstance(e, dict):
                assert r == "raises"
            else:
                assert records_to_dict(q, r) == e


@pytest.mark.parametrize("uniq", [0, 1, 2])
@pytest.mark.parametrize("seed_K", [0, 1, 2, 3, 4, 6, 8])
def test_maximal_matches_are_exactly_the_right_maximal_LS_pairs(emus, seed_K, uniq):
    """sweep output == {(LS[j], j) : j == L or LS[j+1] > LS[j]} with true SA intervals, with and without the seed table
    and the unique-match shortcut (text comparison instead of FM steps once a match occurs once; logic only -- the
    kernel compiles it out, profiles/r01_notes.md)."""
    g, em = emus["medium_data"]
    em.seed_K = seed_K
    em.uniq = uniq
    idx = rp.RefIndex(g["text"], g["suffix_array"])
    rng = random.Random(5)
    for _ in range(150):
        L = rng.randint(1, 90)
        if rng.random() < 0.5:
            p = rng.randrange(0, len(g["text"]) - L)
            q = list(g["text"][p:p + L])
            for k in range(L):
                if rng.random() < 0.05:
                    q[k] = rng.choice("ACGT")
            q = "".join(q)
        else:
            q = "".join(rng.choice("ACGT") for _ in range(L))
        LS = [0] * (L + 2)
        for j in range(1, L + 1):
            i = j - 1
            while i > 0 and idx.exact_match_back_prop(q[i - 1:j]) != -1:
                i -= 1
            LS[j] = i
        exp = []
        for j in range(1, L + 1):
            if j == L or LS[j + 1] > LS[j]:
                lo, hi = idx.exact_match_back_prop(q[LS[j]:j])
                exp.append((LS[j], j, lo, hi - lo + 1))
        got, steps = em.sweep(q)
        assert got == exp
        assert steps <= 6 * L + 40


ACGT = st.text(alphabet="ACGT", min_size=1, max_size=60)


@settings(max_examples=60, deadline=None)
@given(text=st.one_of(st.text(alphabet="ACGT", min_size=8, max_size=200),
                      st.builds(lambda u, k: (u * k)[:200], st.text(alphabet="ACGT", min_size=1, max_size=7), st.integers(2, 40))),
       reads=st.lists(ACGT, min_size=1, max_size=6), K=st.integers(2, 5), seed_K=st.integers(0, 6), uniq=st.integers(0, 2))
def test_property_random_and_repetitive_references(text, reads, K, seed_K, uniq):
    if len(set(text)) < 4:
        text = text + "ACGT"           # parity domain: all four bases occur
    em = Emu(text)
    em.seed_K = seed_K
    em.uniq = uniq
    reads = reads + [text[max(0, len(text) - 40):], text[:50]]     # matches touching both ends of the text
    idx = rp.RefIndex(text)
    o = rp.RefSMEM(idx, lut=rp.RefLUT(idx, K))
    for q in reads:
        assert records_to_dict(q, em.smem(0, q, 1)) == gu.norm(o.get_SMEMS(q, 1))
        if len(q) >= K:
            assert records_to_dict(q, em.smem(1, q, 1, K)) == gu.norm(o.get_smems_lut(q))


def test_long_low_complexity_read_overflows_candidate_cache(emus):
    """> 32 occurrence-count changes in one forward extension: exercises the candidate spill path."""
    text = "A" * 300 + "C" + "A" * 120 + "G" + "ACGT" * 5 + "T" * 80
    em = Emu(text)
    o = rp.RefSMEM(rp.RefIndex(text))
    for seed_K in (0, 3, 7):
        em.seed_K = seed_K
        for q in ["A" * 151, "A" * 100 + "C" + "A" * 50, "T" * 70 + "A" * 81, "A" * 40 + "G" + "ACGT" * 3 + "T" * 60]:
            assert records_to_dict(q, em.smem(0, q, 1)) == gu.norm(o.get_SMEMS(q, 1))
    em.seed_K = 0
    for q in ["A" * 151, "A" * 100 + "C" + "A" * 50, "T" * 70 + "A" * 81, "A" * 40 + "G" + "ACGT" * 3 + "T" * 60]:
        assert records_to_dict(q, em.smem(0, q, 1)) == gu.norm(o.get_SMEMS(q, 1))


def test_rmi_vectorised_fit_matches_loop_fit():
    """The segment-reduction trainer (large models) follows the same rule as the per-bucket loop that is pinned to
    the reference-trained parameters: predictions agree to rounding on every key."""
    import numpy as np
    from genie_smem_b200.surface import RMI
    rng = np.random.default_rng(4)
    keys = np.sort(rng.integers(0, 4 ** 12, 60000)).astype(np.int64)
    keys[100:140] = keys[100]                       # a flat run: zero-variance bucket
    rows = np.arange(len(keys))
    for experts in ([10, 100], [7], [50, 3000]):    # [50, 3000] leaves many leaf buckets empty (root aliasing)
        a = RMI(experts).fit(keys, rows, vectorised=False)
        b = RMI(experts).fit(keys, rows, vectorised=True)
        assert a.level_sizes == b.level_sizes and a.coef.shape == b.coef.shape
        assert np.abs(a.predict(keys) - b.predict(keys)).max() < 1e-6


@pytest.mark.parametrize("name", ["small_data", "medium_data", "big_data"])
def test_lf_walk_to_sampled_rows_recovers_the_suffix_array(emus, name):
    """lf_single + sampled SA (the locate kernel's arithmetic) == the reference's suffix_array, every row."""
    g, em = emus[name]
    sa = np.asarray(g["suffix_array"], np.uint32)
    rows = np.arange(len(sa), dtype=np.uint32) if len(sa) < 5000 else np.random.default_rng(1).integers(0, len(sa), 4000).astype(np.uint32)
    for sample in (1, 2, 7, 32, 1 << 30):
        assert np.array_equal(em.locate(rows, sample), sa[rows]), sample


def _bad_model(p, rng):
    """The golden model with its leaf intercepts shifted by up to a few hundred rows: predictions far off, many hazards."""
    q = dict(p)
    icpt = np.array(p["intercept"], np.float64)
    n_root = int(p["level_sizes"][0])
    icpt[n_root:] += rng.integers(-300, 300, len(icpt) - n_root)
    q["intercept"] = icpt
    return q


@pytest.mark.parametrize("tag,name,bad", [("medium_data_k6", "medium_data", False), ("medium_data_k6", "medium_data", True)])
def test_rmi_hazard_free_reads_take_the_picks(emus, tag, name, bad):
    """The RMI-SMEM pre-filter (k_rmi_hazard_scan / k_rmi_prefilter): (1) a code outside the hazard set looks up exactly -- the
    literal search returns the k-mer's true interval; (2) rmi_read_hazard_free on rolled window codes == membership of every
    kmer_code window in the set; (3) a read without a hazard window emits, through the frame machine (Selector::run_seeded<RMI>),
    exactly the records of BWA-SMEM with min_len 1 (Selector::run_bwa) -- what lets gsm_smem_select(RMI) hand such reads to
    k_select<BWA>."""
    from tests.emu.harness import hazard_table
    g, em = emus[name]
    text = g["text"]
    rng = np.random.default_rng(11)
    p = gu.load_rmi(tag)
    if bad:
        p = _bad_model(p, rng)
    K = p["K"]
    hz = em.rmi_hazards(p)
    hz_set = set(int(c) for c in hz)
    slots = hazard_table(em.lib, hz)
    assert slots is not None and (slots != 0xFFFFFFFF).sum() == len(hz_set)
    assert 0 < len(hz_set) < 4 ** K
    # (1) exact lookups outside the set; inside the set the table says so
    codes = np.arange(4 ** K) if K <= 6 else rng.integers(0, 4 ** K, 4000)
    for code in codes:
        code = int(code)
        inside = bool(em.lib.emu_hz_contains(slots.ctypes.data, len(slots), code))
        assert inside == (code in hz_set)
        hazard, lo, hi = em.rmi_arith_lookup(p, code)
        assert hazard == inside
        if not inside:
            s, _, glo, ghi = em.rmi_lookup(p, code)           # the literal RMI_LUT search
            assert s == 0 and (glo, ghi) == (lo, hi), code
            kmer = "".join("ACGT"[(code >> (2 * (K - 1 - t))) & 3] for t in range(K))
            assert (hi - lo + 1) == sum(1 for i in range(len(text) - K + 1) if text.startswith(kmer, i)) or K > 6
    # (2), (3)
    reads = []
    rnd = random.Random(5)
    for _ in range(220):
        L = rnd.choice((K, K + 1, 25, 60, 101, 151))
        L = min(L, len(text) - 1)
        s0 = rnd.randrange(0, len(text) - L)
        q = list(text[s0:s0 + L])
        pm = rnd.choice((0.0, 0.01, 0.05, 0.3))
        for k in range(L):
            if rnd.random() < pm:
                q[k] = rnd.choice("ACGT")
        reads.append("".join(q))
    reads += ["".join(rnd.choice("ACGT") for _ in range(rnd.choice((K, 40, 101)))) for _ in range(40)]
    reads += ["A" * 40, "AC" * 25, text[-60:], text[:50], text[-(K + 3):]]
    n_free = n_hazard = 0
    em.rmi_fast = True
    try:
        for q in reads:
            w = em.pack_read(q)
            window_codes = [int(em.lib.emu_kmer_code(w.ctypes.data, i, K)) for i in range(len(q) - K + 1)]
            want = not any(c in hz_set for c in window_codes)
            got = bool(em.lib.emu_read_hazard_free(w.ctypes.data, len(q), K, slots.ctypes.data, len(slots)))
            assert got == want, q
            if got:
                n_free += 1
                assert em.smem(2, q, rmi=p) == em.smem(0, q, min_len=1), q
            else:
                n_hazard += 1
    finally:
        em.rmi_fast = False
    assert n_free >= 20 and n_hazard >= 1, (n_free, n_hazard)


@pytest.mark.parametrize("tag,name", [("medium_data_k6", "medium_data"), ("big_data_k12", "big_data"), ("big_data_k15", "big_data")])
def test_get_smems_rmi_equals_get_SMEMS_when_every_window_looks_up_exactly(tag, name):
    """The identity behind the RMI-SMEM pre-filter, on the literal Python restatement (oracle/ref_port.py, pinned to the golden
    vectors): if get_suffix_rmi returns the true interval for every K-mer window of a read (hit <=> the k-mer occurs),
    get_smems_rmi(q) == get_SMEMS(q, 1) as ordered dicts -- with the reference-trained models and with perturbed ones."""
    g = gu.load_index(name)
    text = g["text"]
    idx = rp.RefIndex(text, g["suffix_array"])
    p = gu.load_rmi(tag)
    K = p["K"]
    rnd = random.Random(len(text) + K)
    n_same = n_skipped = 0
    for bad in (False, True):
        icpt = np.array(p["intercept"], np.float64)
        if bad:
            n_root = int(p["level_sizes"][0])
            icpt[n_root:] += np.random.default_rng(3).integers(-40, 40, len(icpt) - n_root)
        rmi = rp.RefRMI(idx, K, p["level_sizes"], p["coef"], icpt)
        o = rp.RefSMEM(idx, rmi=rmi)
        for _ in range(60 if K > 6 else 120):
            L = rnd.choice((K, K + 2, 30, 60, 101))
            s0 = rnd.randrange(0, len(text) - L)
            q = list(text[s0:s0 + L])
            pm = rnd.choice((0.0, 0.02, 0.1))
            for k in range(L):
                if rnd.random() < pm:
                    q[k] = rnd.choice("ACGT")
            q = "".join(q)
            exact = True
            for i in range(L - K + 1):
                kmer = q[i:i + K]
                true = idx.exact_match_back_prop(kmer)
                try:
                    got = rmi.get_suffix_rmi(kmer)
                except (IndexError, RecursionError, TypeError):
                    exact = False
                    break
                if (true == -1 and got[1] >= got[0]) or (true != -1 and tuple(got) != tuple(true)):
                    exact = False
                    break
            if not exact:
                n_skipped += 1
                continue
            a = o.get_SMEMS(q, 1)
            b = {k: tuple(v) for k, v in o.get_smems_rmi(q).items()}
            assert a == b and list(a) == list(b), (bad, q)
            n_same += 1
    assert n_same >= 40, (n_same, n_skipped)


@settings(max_examples=150, deadline=None)
@given(st.text(alphabet="ACGT", min_size=0, max_size=200), st.integers(1, 15), st.integers(0, 2 ** 32 - 1))
def test_read_hazard_free_is_window_membership(q, K, seed):
    """rmi_read_hazard_free (rolled 2-bit window codes + open-addressing probes) == "no window's kmer_code is in the set", for
    any read length (also shorter than K, also ending exactly at a 16-base word boundary), any K <= 15 and sets that hold some
    of the read's own windows."""
    from tests.emu.harness import build, hazard_table
    import ctypes as C
    lib = C.CDLL(build())
    lib.emu_read_hazard_free.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    lib.emu_hz_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    rnd = random.Random(seed)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    windows = []
    for i in range(len(q) - K + 1):
        v = 0
        for ch in q[i:i + K]:
            v = v * 4 + code[ch]
        windows.append(v)
    members = set(rnd.randrange(0, 4 ** K) for _ in range(rnd.choice((0, 1, 5, 300))))
    if windows and rnd.random() < 0.5:
        members |= set(rnd.sample(windows, min(len(windows), rnd.choice((1, 2)))))
    slots = hazard_table(lib, np.asarray(sorted(members), np.uint32))
    assert slots is not None
    words = np.zeros(len(q) // 16 + 3, np.uint32)
    for i, ch in enumerate(q):
        words[i >> 4] |= np.uint32(code[ch] << (30 - 2 * (i & 15)))
    got = bool(lib.emu_read_hazard_free(words.ctypes.data, len(q), K, slots.ctypes.data, len(slots)))
    assert got == (not any(w in members for w in windows))
