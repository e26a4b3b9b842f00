// Per-read selection logic of the three reference SMEM entry points, written once for device and
// for the host-compiled logic test.  Input: the read, and the sorted list of its maximal exact
// matches (sweep_logic.cuh).  Output: the records the reference would put in its result dict,
// in insertion order, before the dict collapses duplicate strings.
//
//   BWA: SMEM.get_SMEMS / get_SMEM_at_index            reference SMEM/SMEM.py:456-484
//   LUT: SMEM.get_smems_lut                            reference SMEM/SMEM.py:20-192
//   RMI: SMEM.get_smems_rmi                            reference SMEM/SMEM.py:206-384
//        + RMI.predict (RMI.py:52-69), RMI_LUT.get_suffix_rmi / exponential_search /
//          binary_search / get_ref_seq (RMI_LUT.py:53-184)
//
// Notation: LS[j] = leftmost start of a match ending at j, F(p) = end of the
// longest match starting at p.  With M = sorted maximal matches (s_k, e_k):
//   F(p)  = e_k of the last k with s_k <= p;      LS[j] = s_k of the first k with e_k >= j.
#pragma once
#include "fm_core.cuh"
#include "sweep_logic.cuh"

namespace gsm {

// ------------------------------------------------------------------------------------ RMI
struct RmiModel {
    uint32_t K;
    uint32_t n_levels;
    uint32_t level_size[8];
    uint32_t level_off[8];
    const double* coef;
    const double* intercept;
};

GSM_HD double mul_add_nofma(double x, double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dmul_rn(x, a), b);
#else
    volatile double t = x * a;   // the host test is also built with -ffp-contract=off
    return t + b;
#endif
}

// RMI.predict for one key (RMI.py:52-69): per level p = fl(fl(x*coef)+intercept); route with
// min(scale-1, max(0, int(p))), scale = size of the next level (experts + [1]).
GSM_HD double rmi_predict(const RmiModel& m, uint64_t code) {
    const double x = (double)code;
    uint32_t model = 0;
    double p = 0.0;
    for (uint32_t lv = 0; lv < m.n_levels; ++lv) {
        const uint32_t k = m.level_off[lv] + model;
        p = mul_add_nofma(x, m.coef[k], m.intercept[k]);
        const uint32_t scale = (lv + 1 < m.n_levels) ? m.level_size[lv + 1] : 1u;
        if (!(p >= 1.0)) model = 0;                       // int(p) <= 0 (also NaN)
        else if (p >= (double)scale) model = scale - 1;
        else model = (uint32_t)p;
    }
    return p;
}

// Table access for the last-mile search.  Probe(row, s, code64): s = suffix_array[row] (1-based) and
// code64 = the MSB-first code of the 32 bases at text position s-1 (zero-padded past the end), either
// from the suffix array + packed text (two dependent fetches) or from a precomputed 16-byte probe
// record {s, code64} (one fetch, gsm_rmi_probe_build).
template <typename SaLoad, typename TextLoad>
struct SaTextProbe {
    SaLoad sa;
    TextLoad text;
    GSM_HD void operator()(uint64_t row, int64_t& s, uint64_t& code64) const {
        s = (int64_t)sa(row);
        code64 = kmer_code(text, (uint64_t)(s - 1), 32);
    }
};

template <typename Probe>
struct RmiTable {
    Probe probe;
    int64_t n_rows;     // len(suffix_array)
    int64_t n_bases;    // ref_seq_size
    uint32_t K;
    bool raised;        // the reference would raise IndexError / RecursionError

    // RMI_LUT.get_ref_seq (RMI_LUT.py:89-92) with Python list indexing: rows in [-n, n) are
    // valid (negative wraps), anything else raises IndexError.  Returns false for None.
    GSM_HD bool ref_seq(int64_t ind, uint64_t& code) {
        code = 0;
        if (ind < -n_rows || ind >= n_rows) { raised = true; return false; }
        if (ind < 0) ind += n_rows;
        int64_t s;
        uint64_t code64;
        probe((uint64_t)ind, s, code64);
        if (s - 1 + (int64_t)K > n_bases) return false;
        code = code64 >> (64u - 2u * K);
        return true;
    }

    // RMI_LUT.binary_search (RMI_LUT.py:95-133), iterative; the recursion is a tail call.
    GSM_HD int64_t binary_search(uint64_t q, int64_t lower, int64_t upper, bool strict) {
        for (int depth = 0; depth < 400; ++depth) {
            if (raised) return 0;
            if (lower == upper) return lower;
            uint64_t c;
            bool ok;
            if (upper - lower == 1) {
                if (strict) { ok = ref_seq(upper, c); return (ok && c == q) ? upper : lower; }
                ok = ref_seq(lower, c);
                return (ok && c == q) ? lower : upper;
            }
            // Python floor division
            int64_t sum = lower + upper;
            int64_t mid = (sum >= 0) ? sum / 2 : -((-sum + 1) / 2);
            uint64_t mc;
            bool mok = ref_seq(mid, mc);
            while (!mok && mid > lower && !raised) {
                mid -= 1;
                mok = ref_seq(mid, mc);
                if (mid == lower) {
                    if (strict) { ok = ref_seq(upper, c); return (ok && c == q) ? upper : lower; }
                    return (mok && mc == q) ? lower : upper;
                }
            }
            if (raised) return 0;
            if (!mok) { raised = true; return 0; }       // None < str: TypeError in the reference
            if (mc < q || (mc == q && strict)) lower = mid; else upper = mid;
        }
        raised = true;   // non-terminating recursion: RecursionError in the reference
        return 0;
    }

    // RMI_LUT.exponential_search (RMI_LUT.py:136-184)
    GSM_HD void exponential_search(uint64_t q, int64_t start, int64_t& out_lo, int64_t& out_hi) {
        out_lo = 0; out_hi = -1;
        bool have_lower = false, have_upper = false;
        int64_t lower = 0, upper = 0;
        uint64_t cur;
        bool ok = ref_seq(start, cur);
        while (!ok && !raised) { start += 1; ok = ref_seq(start, cur); }
        if (raised) return;
        if (cur < q) { lower = start; have_lower = true; }
        else if (cur > q) { upper = start; have_upper = true; }
        int64_t win = 1;
        if (!have_upper) {
            while (start + win < n_bases + 1) {
                int64_t ind = start + win;
                win *= 2;
                uint64_t f;
                bool fok = ref_seq(ind, f);
                while (!fok && !raised) { ind += 1; fok = ref_seq(ind, f); }
                if (raised) return;
                if (f > q) { upper = ind; have_upper = true; break; }
                if (f < q) { lower = ind; have_lower = true; }
            }
        }
        win = 1;
        if (!have_lower) {
            while (start - win >= 0) {
                int64_t ind = start - win;
                win *= 2;
                uint64_t f;
                bool fok = ref_seq(ind, f);
                while (!fok && !raised) { ind -= 1; fok = ref_seq(ind, f); }
                if (raised) return;
                if (f < q) { lower = ind; have_lower = true; break; }
                if (f > q) { upper = ind; have_upper = true; }
            }
        }
        if (!have_lower) lower = 0;
        if (!have_upper) upper = n_rows - 1;
        // The two binary searches of RMI_LUT.py:183-184 start from the same bracket and take the same
        // branch as long as the probed k-mer differs from q and no None row touches `lower`: walk that
        // shared prefix once (same probes, same order), then finish each search on its own.
        for (int depth = 0; depth < 400 && upper - lower > 1; ++depth) {
            const int64_t sum = lower + upper;
            int64_t mid = (sum >= 0) ? sum / 2 : -((-sum + 1) / 2);
            uint64_t mc;
            bool mok = ref_seq(mid, mc);
            bool at_lower = false;
            while (!mok && mid > lower && !raised) {
                mid -= 1;
                mok = ref_seq(mid, mc);
                if (mid == lower) { at_lower = true; break; }
            }
            if (raised) return;
            if (at_lower || !mok || mc == q) break;       // the searches part ways here
            if (mc < q) lower = mid; else upper = mid;
        }
        out_lo = binary_search(q, lower, upper, false);
        if (raised) return;
        out_hi = binary_search(q, lower, upper, true);
    }

    // RMI_LUT.get_suffix_rmi (RMI_LUT.py:67-78): int(prediction) truncates toward zero.
    GSM_HD void lookup(const RmiModel& m, uint64_t q, double& pred, int64_t& lo, int64_t& hi) {
        pred = rmi_predict(m, q);
        lo = 0; hi = -1;
        if (!(pred > -9.0e18 && pred < 9.0e18)) { raised = true; return; }
        exponential_search(q, (int64_t)pred, lo, hi);
    }
};

// ------------------------------------------------------------------------------------ selection
// Ctx must provide:
//   uint32_t L, K, n_mems;  uint32_t min_len;
//   MemEntry mem(uint32_t k)                 k-th maximal match, sorted by end (and start)
//   uint32_t base(uint32_t pos)
//   bool seed(uint32_t c, int64_t& lo, int64_t& hi)      LUT / RMI lookup of q[c:c+K]; true on a hit
//   bool sequential(uint32_t c, int64_t clo, int64_t chi, uint32_t pc, int64_t plo, int64_t phi)
//                                            check_sequential of the two seeds (SMEM.py:196-202)
//   void interval(uint32_t i, uint32_t j, uint32_t& lo, uint32_t& cnt)   true SA interval of q[i:j]
//   void emit(uint32_t i, uint32_t j, int64_t lo, int64_t hi)
//   bool failed()                             the reference raised inside seed()
//   bool seeds_are_true()                     seed() returns exact SA intervals (LUT), not RMI guesses
template <typename Ctx>
struct Selector {
    GSM_HD static uint32_t s_of(const MemEntry& e) { return e.se & 0xFFFFu; }
    GSM_HD static uint32_t e_of(const MemEntry& e) { return e.se >> 16; }

    // get_SMEM_at_index (SMEM.py:469-484) == longest maximal match covering p, ties to the
    // smallest end (the strict '>' of SMEM.py:413 scanning ends upward).  `from` = first k with
    // e_k > p.  Returns the index of the winner.
    GSM_HD static uint32_t covering_best(Ctx& c, uint32_t p, uint32_t& from) {
        while (from < c.n_mems && e_of(c.mem(from)) <= p) ++from;
        uint32_t best = from, bestlen = 0;
        for (uint32_t k = from; k < c.n_mems; ++k) {
            MemEntry m = c.mem(k);
            if (s_of(m) > p) break;
            uint32_t len = e_of(m) - s_of(m);
            if (len > bestlen) { bestlen = len; best = k; }
        }
        return best;
    }

    // get_SMEMS (SMEM.py:456-467)
    GSM_HD static void run_bwa(Ctx& c) {
        uint32_t p = 0, from = 0;
        while (p < c.L && from < c.n_mems) {
            uint32_t b = covering_best(c, p, from);
            if (b >= c.n_mems) break;
            MemEntry m = c.mem(b);
            if (e_of(m) - s_of(m) >= c.min_len) c.emit(s_of(m), e_of(m), (int64_t)m.lo, (int64_t)m.lo + m.cnt - 1);
            p = e_of(m);
        }
    }

    struct Cand {
        bool valid;
        uint32_t i, j;
        int64_t lo, hi;
    };

    GSM_HD static void upd(Cand& cd, uint32_t i, uint32_t j, int64_t lo, int64_t hi) {
        if (!cd.valid || (j - i) >= (cd.j - cd.i)) { cd.valid = true; cd.i = i; cd.j = j; cd.lo = lo; cd.hi = hi; }
    }

    // F(p) restricted to true matches; 0 matches => p (cannot happen when all four bases occur)
    GSM_HD static uint32_t F_of(Ctx& c, uint32_t p) {
        uint32_t f = p;
        for (uint32_t k = 0; k < c.n_mems; ++k) {
            MemEntry m = c.mem(k);
            if (s_of(m) > p) break;
            if (e_of(m) > f) f = e_of(m);
        }
        return f;
    }

    // True SA interval of q[i:j): read it off the match list when (i, j) is itself a maximal match
    // (the usual case), otherwise one backward search from j down to i.
    GSM_HD static void true_iv(Ctx& c, uint32_t i, uint32_t j, uint32_t& lo, uint32_t& cnt) {
        for (uint32_t k = 0; k < c.n_mems; ++k) {
            MemEntry m = c.mem(k);
            if (e_of(m) < j) continue;
            if (e_of(m) == j && s_of(m) == i) { lo = m.lo; cnt = m.cnt; return; }
            break;
        }
        c.interval(i, j, lo, cnt);
    }

    // forward_extension(query, pc+K, kmer, seed) (SMEM.py:425-443): longest key and its value
    GSM_HD static void fwd_only(Ctx& c, uint32_t pc, int64_t slo, int64_t shi, uint32_t& end, int64_t& lo, int64_t& hi) {
        uint32_t f = F_of(c, pc);
        if (f <= pc + c.K) { end = pc + c.K; lo = slo; hi = shi; return; }   // the seed key itself
        end = f;
        uint32_t l, n;
        true_iv(c, pc, f, l, n);
        lo = (int64_t)l; hi = (int64_t)l + n - 1;
    }

    // backward_extension(query, pc, keys) (SMEM.py:389-423) over keys pc+K .. max(F(pc), pc+K)
    // (all_keys) or over the seed key only.
    GSM_HD static void bext(Ctx& c, uint32_t pc, int64_t slo, int64_t shi, bool all_keys, Cand& out) {
        const uint32_t K = c.K;
        uint32_t f = F_of(c, pc);
        const bool seed_true = f >= pc + K;           // the k-mer really occurs
        uint32_t jmax = all_keys ? (f > pc + K ? f : pc + K) : pc + K;
        // extended keys: j in [pc+K, jmax] with LS[j] < pc; only maximal-match ends and jmax matter
        bool have = false;
        uint32_t bi = 0, bj = 0, blo = 0, bcnt = 0;
        bool b_from_mem = false;
        if (seed_true) {
            for (uint32_t k = 0; k < c.n_mems; ++k) {
                MemEntry m = c.mem(k);
                uint32_t s = s_of(m), e = e_of(m);
                if (e < pc + K) continue;
                if (s >= pc) break;                    // starts are sorted: no further left extension
                uint32_t j = e <= jmax ? e : jmax;     // plateau end, or the key range's last key
                if (!have || (j - s) > (bj - bi)) {
                    have = true; bi = s; bj = j;
                    b_from_mem = (j == e);
                    blo = m.lo; bcnt = m.cnt;
                }
                if (e >= jmax) break;
            }
        }
        // the longest key wins only if strictly longer (SMEM.py:418)
        uint32_t fend = jmax;
        if (!have || (fend - pc) > (bj - bi)) {
            out.valid = true; out.i = pc; out.j = fend;
            if (fend == pc + K) { out.lo = slo; out.hi = shi; }
            else { uint32_t l, n; true_iv(c, pc, fend, l, n); out.lo = (int64_t)l; out.hi = (int64_t)l + n - 1; }
            return;
        }
        out.valid = true; out.i = bi; out.j = bj;
        if (!b_from_mem) true_iv(c, bi, bj, blo, bcnt);
        out.lo = (int64_t)blo; out.hi = (int64_t)blo + bcnt - 1;
    }

    // get_smems_lut / get_smems_rmi: the frame machine of SMEM.py:49-186 / :235-379.
    GSM_HD static void run_seeded(Ctx& c) {
        const uint32_t K = c.K, L = c.L;
        bool first = true;
        uint32_t e = 0, plen = 0;
        for (;;) {
            if (!first && e >= L) return;
            // frame: 0 = None, 1 = () , 2 = k-mer frame
            int fstate = 0;
            uint32_t pc = 0; bool pfw = false; int64_t plo = 0, phi = -1;
            Cand cd; cd.valid = false; cd.i = cd.j = 0; cd.lo = 0; cd.hi = -1;
            const uint32_t pstart = e - plen;
            const uint32_t nwin = first ? 1u : K;
            for (uint32_t i = 0; i < nwin; ++i) {
                uint32_t cpos;
                if (first) cpos = 0;
                else {
                    if (i >= plen) continue;
                    cpos = e - i;
                    if (cpos + K > L) continue;
                }
                int64_t lo, hi;
                const bool hit = c.seed(cpos, lo, hi);        // the ONLY lookup site
                if (c.failed()) return;
                if (first) {                                   // SMEM.py:26-39 / :213-225
                    uint32_t end; int64_t flo, fhi;
                    if (hit) fwd_only(c, 0, lo, hi, end, flo, fhi);
                    else {
                        end = F_of(c, 0);
                        uint32_t l, n; true_iv(c, 0, end, l, n); flo = (int64_t)l; fhi = (int64_t)l + n - 1;
                    }
                    c.emit(0, end, flo, fhi);
                    e = end; plen = end;
                    break;
                }
                if (hit) {
                    if (fstate == 0) { fstate = 2; pc = cpos; pfw = true; plo = lo; phi = hi; }          // :70
                    else if (fstate == 1) { fstate = 2; pc = cpos; pfw = false; plo = lo; phi = hi; }    // :73
                    else {
                        // check_sequential of two TRUE k-mer intervals at adjacent windows is just
                        // "q[cpos : cpos+K+1) occurs" (SURVEY A13): read it off the match list
                        const bool seq = (c.seeds_are_true() && pc == cpos + 1) ? (F_of(c, cpos) >= cpos + K + 1)
                                                                               : c.sequential(cpos, lo, hi, pc, plo, phi);
                        if (seq) {                                                                        // Case 1
                            if (pfw) { Cand b; bext(c, pc, plo, phi, true, b); upd(cd, b.i, b.j, b.lo, b.hi); }
                            else {
                                if (cd.valid && (pc - pstart) + K < (cd.j - cd.i)) continue;              // :94-95
                                Cand b; bext(c, pc, plo, phi, false, b); upd(cd, b.i, b.j, b.lo, b.hi);
                            }
                        } else {                                                                          // Case 2
                            if (pfw) { uint32_t end; int64_t flo, fhi; fwd_only(c, pc, plo, phi, end, flo, fhi); upd(cd, pc, end, flo, fhi); }
                            else upd(cd, cpos, cpos + K, lo, hi);
                        }
                        pc = cpos; pfw = false; plo = lo; phi = hi;
                    }
                } else {
                    if (fstate == 2) {                                                                    // Case 3
                        if (pfw) { uint32_t end; int64_t flo, fhi; fwd_only(c, pc, plo, phi, end, flo, fhi); upd(cd, pc, end, flo, fhi); }
                        else upd(cd, pc, pc + K, plo, phi);
                    }
                    fstate = 1;
                }
            }
            if (first) { first = false; continue; }
            if (fstate == 2) {                                                                            // :149-171
                Cand b; bext(c, pc, plo, phi, pfw, b); upd(cd, b.i, b.j, b.lo, b.hi);
            }
            if (!cd.valid) {                                                                              // :175-179
                uint32_t from = 0;
                uint32_t b = covering_best(c, e, from);
                if (b >= c.n_mems) return;
                MemEntry m = c.mem(b);
                c.emit(s_of(m), e_of(m), (int64_t)m.lo, (int64_t)m.lo + m.cnt - 1);
                plen = e_of(m) - s_of(m);
                e = e_of(m);
            } else {
                c.emit(cd.i, cd.j, cd.lo, cd.hi);
                e = cd.j; plen = cd.j - cd.i;
            }
        }
    }
};

}  // namespace gsm
