"""CPU restatement of GENIE-SMEM's search path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (genie_smem_b200) never does.

Parity status: PINNED.  Every routine here is checked in tests/test_oracle_golden.py against
fixtures produced by running the unmodified reference (tests/golden/make_golden.py): index
arrays of mississippi/small/medium/big_data, the reference's checked-in medium_data K=6 LUT,
the paper's mississippi known answers, and frozen outputs of get_SMEMS / get_smems_lut /
get_smems_rmi / get_suffix_rmi on seeded read sets.

Each function cites the reference file:line (under /root/reference/) it restates.  The
restatement is deliberately literal -- same string-keyed dicts, same O(L^2) restarts, same
quirks -- so that it can also serve as the "port" CPU baseline.
"""
from __future__ import annotations

import numpy as np

BASES = "ACGT"
CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


# --------------------------------------------------------------------------- index arrays
def suffix_array(codes: np.ndarray) -> np.ndarray:
    """0-based suffix start positions of `codes` in lexicographic order (prefix doubling).

    `codes` must end with a unique smallest symbol (the '$').  Equivalent to sorting all
    rotations as SMEM/ExactMatch.py:52-58 does, because the unique '$' decides every
    comparison before it wraps.
    """
    n = len(codes)
    rank = np.asarray(codes, dtype=np.int64)
    sa = np.argsort(rank, kind="stable")
    k = 1
    while True:
        r1 = rank
        r2 = np.full(n, -1, dtype=np.int64)
        r2[: n - k] = rank[k:]
        sa = np.lexsort((r2, r1))
        a, b = r1[sa], r2[sa]
        diff = np.empty(n, dtype=np.int64)
        diff[0] = 0
        diff[1:] = (a[1:] != a[:-1]) | (b[1:] != b[:-1])
        newrank = np.cumsum(diff)
        rank = np.empty(n, dtype=np.int64)
        rank[sa] = newrank
        if newrank[-1] == n - 1:
            return sa
        k *= 2


class RefIndex:
    """The `fm_index` dict of SMEM/ExactMatch.py:29-30 as numpy arrays (same values).

    suffix_array : 1-based suffix starts, SA[0] = ref_size       (ExactMatch.py:66)
    bwt          : last column, as a str                           (ExactMatch.py:64)
    occ[c][i]    : INCLUSIVE count of c in bwt[0..i]               (ExactMatch.py:70-90)
    count_dic[c] : first row whose first char is c, "" -> n        (ExactMatch.py:92-101)
    """

    def __init__(self, text: str, suffix_array_1based: np.ndarray | None = None):
        self.ref = text                      # without '$' (LUT.py:18, RMI_LUT.py:24-25)
        self.ref_sequence = text + "$"       # ExactMatch.py:49
        self.ref_size = len(text) + 1        # ExactMatch.py:50
        n = self.ref_size
        alphabet = sorted(set(self.ref_sequence))
        assert alphabet[0] == "$"
        cmap = {c: i for i, c in enumerate(alphabet)}
        codes = np.fromiter((cmap[c] for c in self.ref_sequence), dtype=np.int64, count=n)
        if suffix_array_1based is None:
            sa0 = suffix_array(codes)
        else:
            sa0 = np.asarray(suffix_array_1based, dtype=np.int64) - 1
        self.suffix_array = (sa0 + 1).astype(np.int64)
        bwt_codes = codes[(sa0 - 1) % n]
        self.bwt = "".join(alphabet[c] for c in bwt_codes)
        self.occ = {}
        for c in alphabet:
            self.occ[c] = np.cumsum(bwt_codes == cmap[c]).astype(np.int64)
        first = codes[sa0]
        self.count_dic = {}
        for c in alphabet:
            self.count_dic[c] = int(np.searchsorted(first, cmap[c], side="left"))
        self.count_dic[""] = n

    # SMEM/ExactMatch.py:132-151
    def exact_match_back_prop(self, query_seq: str):
        start = 1
        end = self.count_dic[""]
        for ch in reversed(query_seq):
            c0 = self.count_dic[ch]          # KeyError for a char absent from the text, as in the reference
            if start - 1 <= 0:
                start = c0 + 1
            else:
                start = c0 + 1 + int(self.occ[ch][start - 2])
            end = c0 + int(self.occ[ch][end - 1])
            if start > end:
                return -1
        return start - 1, end - 1

    # SMEM/ExactMatch.py:155-171
    def exact_match_back_prop_add_one(self, ch: str, prev):
        start = prev[0] + 1
        end = prev[1] + 1
        c0 = self.count_dic[ch]
        if start - 1 <= 0:
            start = c0 + 1
        else:
            start = c0 + 1 + int(self.occ[ch][start - 2])
        end = c0 + int(self.occ[ch][end - 1])
        if start > end:
            return -1
        return start - 1, end - 1

    # SMEM/ExactMatch.py:174-188
    def exact_match(self, query_seq: str):
        start, end = self.exact_match_back_prop(query_seq)   # TypeError on a miss, as in the reference
        return sorted(int(x) for x in self.suffix_array[start:end + 1])

    # SMEM/ExactMatch.py:191-199
    def get_position(self, i):
        return int(self.suffix_array[i])

    def get_positions(self, lo, hi):
        return [int(self.suffix_array[i]) for i in range(lo, hi + 1)]   # negative rows wrap, as in the reference


# --------------------------------------------------------------------------- LUT (SMEM/LUT.py)
def convert_seq_to_num(seq: str) -> int:
    """SMEM/LUT.py:37-48 -- MSB-first base-4 code, A0 C1 G2 T3."""
    v = 0
    for ch in seq:
        v = (v << 2) | CODE[ch]
    return v


class RefLUT:
    """What LUT.generate_lut (SMEM/LUT.py:15-35) stores, computed on demand.

    The reference dict holds exactly the K-mers that occur in the text, keyed by str(code),
    with value [A2(kmer), get_positions(A2(kmer))].  Membership <=> backward search hits.
    """

    def __init__(self, index: RefIndex, K: int):
        self.index = index
        self.lut_size = K
        self._cache = {}

    def lookup(self, kmer: str):
        """-> None on a miss, else ((lo, hi), positions)."""
        key = convert_seq_to_num(kmer)
        hit = self._cache.get(key, 0)
        if hit == 0:
            t = self.index.exact_match_back_prop(kmer)
            hit = None if t == -1 else (t, self.index.get_positions(t[0], t[1]))
            self._cache[key] = hit
        return hit

    def materialise(self):
        """The full table as {code: (lo, hi)} -- same key set as LUT.py:20-35."""
        K, ref = self.lut_size, self.index.ref
        out = {}
        for p in range(len(ref) - K + 1):
            kmer = ref[p:p + K]
            code = convert_seq_to_num(kmer)
            if code not in out:
                out[code] = self.index.exact_match_back_prop(kmer)
        return out


# --------------------------------------------------------------------------- RMI (SMEM/RMI.py, SMEM/RMI_LUT.py)
class RefRMI:
    """RMI.predict (SMEM/RMI.py:52-69) + RMI_LUT lookup (SMEM/RMI_LUT.py:53-184).

    Parameters come from a trained reference model (coef_/intercept_ per LinearRegression);
    sklearn's predict is bit-identical to fl(fl(x*coef)+intercept) in float64.
    level_sizes = number of models per level, e.g. [1, 10, 100] for experts [10, 100].
    """

    def __init__(self, index: RefIndex, K: int, level_sizes, coef, intercept):
        self.index = index
        self.prediction_size = K
        self.level_sizes = [int(x) for x in level_sizes]
        self.coef = np.asarray(coef, dtype=np.float64)
        self.intercept = np.asarray(intercept, dtype=np.float64)
        self.level_off = np.concatenate([[0], np.cumsum(self.level_sizes)])[:-1]
        # routing scales: experts + [1]  (RMI.py:54); level l routes into level_sizes[l+1] models
        self.scales = self.level_sizes[1:] + [1]
        self.ref_seq = index.ref
        self.ref_seq_size = index.ref_size - 1      # RMI_LUT.py:31
        self.sa = index.suffix_array

    # RMI.py:52-69 for a single key
    def rmi_predict(self, kmer: str) -> float:
        x = np.float64(convert_seq_to_num(kmer[: self.prediction_size]))   # RMI_LUT.py:58-61
        model = 0
        p = np.float64(0.0)
        for lvl, scale in enumerate(self.scales):
            k = self.level_off[lvl] + model
            p = np.float64(x * self.coef[k]) + self.intercept[k]
            model = min(scale - 1, max(0, int(p)))                         # RMI.py:66
        return float(p)

    # RMI_LUT.py:89-92 (Python negative indexing included)
    def get_ref_seq(self, ind: int):
        s = int(self.sa[ind])
        if s - 1 + self.prediction_size > self.ref_seq_size:
            return None
        return self.ref_seq[s - 1: s - 1 + self.prediction_size]

    # RMI_LUT.py:95-133
    def binary_search(self, q, lower, upper, strict):
        if lower == upper:
            return lower
        if upper - lower == 1:
            if strict:
                return upper if self.get_ref_seq(upper) == q else lower
            return lower if self.get_ref_seq(lower) == q else upper
        mid = (lower + upper) // 2
        mid_seq = self.get_ref_seq(mid)
        while mid_seq is None and mid > lower:
            mid -= 1
            mid_seq = self.get_ref_seq(mid)
            if mid == lower:
                if strict:
                    return upper if self.get_ref_seq(upper) == q else lower
                return lower if mid_seq == q else upper
        if mid_seq < q or (mid_seq == q and strict):     # TypeError if mid_seq is None, as in the reference
            return self.binary_search(q, mid, upper, strict)
        return self.binary_search(q, lower, mid, strict)

    # RMI_LUT.py:136-184
    def exponential_search(self, q, start_sa):
        n_rows = len(self.ref_seq) + 1
        lower = upper = None
        cur = self.get_ref_seq(start_sa)
        while cur is None:
            start_sa += 1
            cur = self.get_ref_seq(start_sa)
        if cur < q:
            lower = start_sa
        elif cur > q:
            upper = start_sa
        w = 1
        if upper is None:
            while start_sa + w < n_rows:
                ind = start_sa + w
                w *= 2
                found = self.get_ref_seq(ind)
                while found is None:
                    ind += 1
                    found = self.get_ref_seq(ind)
                if found > q:
                    upper = ind
                    break
                if found < q:
                    lower = ind
        w = 1
        if lower is None:
            while start_sa - w >= 0:
                ind = start_sa - w
                w *= 2
                found = self.get_ref_seq(ind)
                while found is None:
                    ind -= 1
                    found = self.get_ref_seq(ind)
                if found < q:
                    lower = ind
                    break
                if found > q:
                    upper = ind
        if lower is None:
            lower = 0
        if upper is None:
            upper = len(self.sa) - 1
        return (self.binary_search(q, lower, upper, False), self.binary_search(q, lower, upper, True))

    # RMI_LUT.py:67-78
    def get_suffix_rmi(self, kmer: str):
        return self.exponential_search(kmer, int(self.rmi_predict(kmer)))


# --------------------------------------------------------------------------- SMEM (SMEM/SMEM.py)
class RefSMEM:
    def __init__(self, index: RefIndex, lut: RefLUT | None = None, rmi: RefRMI | None = None):
        self.matcher = index
        self.lut = lut
        self.rmi_lut = rmi

    # SMEM.py:16-17
    def get_suffix_index(self, q):
        return self.matcher.exact_match_back_prop(q)

    # SMEM.py:425-443 -- every prefix is searched from scratch (this is the O(L^2) part)
    def forward_extension(self, query, start_index, largest="", suffix_tuple=None):
        found = {}
        if suffix_tuple is not None:
            found[largest] = suffix_tuple
        cur = largest
        for i in range(start_index + 1, len(query) + 1):
            cur = largest + query[start_index:i]
            t = self.get_suffix_index(cur)
            if t == -1:
                return found, cur[:-1]
            found[cur] = t
        return found, cur

    # SMEM.py:389-423
    def backward_extension(self, query, start_index, forward_matches):
        best, best_t, best_end = "", None, -1
        longest_fwd = ""
        for key in forward_matches:
            if len(key) > len(longest_fwd):
                longest_fwd = key
            t = None
            for i in range(start_index - 1, -1, -1):
                cur = query[i:start_index] + key
                if t is None:
                    t = self.get_suffix_index(cur)                               # :406
                else:
                    t = self.matcher.exact_match_back_prop_add_one(cur[0], t)   # :408
                if t == -1:
                    break
                if len(cur) > len(best):                                         # strict, :413
                    best, best_t, best_end = cur, t, start_index + len(key)
        if len(longest_fwd) > len(best):                                         # strict, :418
            best, best_t, best_end = longest_fwd, forward_matches[longest_fwd], start_index + len(longest_fwd)
        return best, best_t, best_end

    # SMEM.py:469-484
    def get_SMEM_at_index(self, query, start_index):
        fwd = self.forward_extension(query, start_index)
        back = self.backward_extension(query, start_index, fwd[0])
        if len(fwd[1]) > len(back[0]):
            return [fwd[1], fwd[0][fwd[1]], len(fwd[1]) + start_index]
        return [back[0], back[1], back[2]]

    # SMEM.py:456-467 -- BWA-SMEM entry point
    def get_SMEMS(self, query, minimum_length):
        p, out = 0, {}
        while p < len(query):
            s = self.get_SMEM_at_index(query, p)
            if len(s[0]) >= minimum_length:
                out[s[0]] = s[1]
            p = s[2]
        return out

    # SMEM.py:196-202
    @staticmethod
    def check_sequential(a, b):
        bs = set(b)
        return any(x + 1 in bs for x in a)

    # SMEM.py:20-192 (LUT) and SMEM.py:206-384 (RMI) are the same machine with a different
    # seed provider; `seed(kmer)` returns None on a miss, else (tuple, positions).
    def _seeded(self, query, K, seed):
        out = {}
        L = len(query)
        first = seed(query[:K])
        if first is not None:                                                    # :31-32 / :217-218
            fm = self.forward_extension(query, K, query[:K], first[0])
        else:                                                                    # :35-36 / :221-222
            fm = self.forward_extension(query, 0)
        out[fm[1]] = fm[0][fm[1]]                                                # :38-39
        prev_len = len(fm[1])
        e = prev_len
        while e < L:                                                             # :49
            prev = None            # None / () / (kmer, tuple, start, fwd, positions)
            cand, cand_t, cand_end = None, (), -1
            prev_start = e - prev_len

            def better(s):
                return cand is None or len(s) >= len(cand)

            for i in range(K):                                                   # :56
                if i >= prev_len:
                    continue
                c = e - i
                if c + K > L:
                    continue
                sub = query[c:c + K]
                hit = seed(sub)
                if hit is not None:
                    if prev is None:
                        prev = (sub, hit[0], c, True, hit[1])                    # :70
                    elif prev == ():
                        prev = (sub, hit[0], c, False, hit[1])                   # :73
                    else:
                        pk, pt, pc, pfw, ppos = prev
                        if self.check_sequential(hit[1], ppos):                  # Case 1, :75
                            if pfw:                                              # :77-84
                                fm = self.forward_extension(query, pc + K, pk, pt)
                                b = self.backward_extension(query, pc, fm[0])
                                if better(b[0]):
                                    cand, cand_t, cand_end = b
                            else:                                                # :92-101
                                if cand is not None and (pc - prev_start) + K < len(cand):
                                    continue                                     # stale frame kept, :94-95
                                b = self.backward_extension(query, pc, {pk: pt})
                                if better(b[0]):
                                    cand, cand_t, cand_end = b
                        else:                                                    # Case 2, :108-122
                            if pfw:
                                fm = self.forward_extension(query, pc + K, pk, pt)
                                if fm[1] != "" and better(fm[1]):
                                    cand, cand_t, cand_end = fm[1], fm[0][fm[1]], len(fm[1]) + pc
                            elif better(sub):
                                cand, cand_t, cand_end = sub, hit[0], K + c
                        prev = (sub, hit[0], c, False, hit[1])                   # :106 / :124
                else:
                    if prev is None or prev == ():
                        prev = ()                                                # :126-128
                    else:                                                        # Case 3, :129-146
                        pk, pt, pc, pfw, ppos = prev
                        if pfw:
                            fm = self.forward_extension(query, pc + K, pk, pt)
                            if fm[1] != "" and better(fm[1]):
                                cand, cand_t, cand_end = fm[1], fm[0][fm[1]], len(fm[1]) + pc
                        elif better(pk):
                            cand, cand_t, cand_end = pk, pt, K + pc
                        prev = ()
            if prev is not None and prev != ():                                  # last frame, :149-171
                pk, pt, pc, pfw, ppos = prev
                if pfw:
                    fm = self.forward_extension(query, pc + K, pk, pt)
                    b = self.backward_extension(query, pc, fm[0])
                else:
                    b = self.backward_extension(query, pc, {pk: pt})
                if better(b[0]):
                    cand, cand_t, cand_end = b
            if cand is None:                                                     # :175-179
                s = self.get_SMEM_at_index(query, e)
                out[s[0]] = s[1]
                e = s[2]
                prev_len = len(s[0])
            else:                                                                # :183-186
                e = cand_end
                out[cand] = cand_t
                prev_len = len(cand)
        return out

    # SMEM.py:20-192 -- LUT-SMEM entry point
    def get_smems_lut(self, query):
        return self._seeded(query, self.lut.lut_size, self.lut.lookup)

    # SMEM.py:206-384 -- RMI-SMEM entry point (the per-call pickle reload of :207 is hoisted)
    def get_smems_rmi(self, query):
        rmi, idx = self.rmi_lut, self.matcher

        def seed(kmer):
            lo, hi = rmi.get_suffix_rmi(kmer)
            if hi >= lo:                                                         # :217, :253
                return (lo, hi), idx.get_positions(lo, hi)                       # :262-263
            return None

        return self._seeded(query, rmi.prediction_size, seed)
