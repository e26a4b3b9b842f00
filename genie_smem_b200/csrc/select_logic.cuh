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
    uint32_t stride;            // distance between consecutive models in coef[] / intercept[] (1 = two dense arrays,
                                // 2 = one interleaved {coef, intercept} array: one 16-byte line per model)
    uint32_t n_none;            // > 0: none_rows holds the rows of the K short suffixes (enables the error-bounded search)
    uint32_t none_rows[32];
    uint32_t none_shift;        // none_map bit (row >> none_shift) is set iff that region of rows holds a None row
    uint32_t none_map[32];      // 1024 regions: the usual bracket is cleared with two bit tests
};

// fill the None-row fields of a model (host side): rows = the K rows of gsm_rmi_none_rows, ascending
inline void rmi_set_none_rows(RmiModel& m, const uint32_t* rows, uint32_t n, uint64_t n_rows) {
    m.n_none = n;
    m.none_shift = 0;
    while ((n_rows >> m.none_shift) > 1024u) ++m.none_shift;
    for (int w = 0; w < 32; ++w) m.none_map[w] = 0;
    for (uint32_t t = 0; t < n && t < 32u; ++t) {
        m.none_rows[t] = rows[t];
        const uint32_t r = rows[t] >> m.none_shift;
        m.none_map[r >> 5] |= 1u << (r & 31u);
    }
}

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
        p = mul_add_nofma(x, m.coef[(size_t)k * m.stride], m.intercept[(size_t)k * m.stride]);
        const uint32_t scale = (lv + 1 < m.n_levels) ? m.level_size[lv + 1] : 1u;
        if (!(p >= 1.0)) model = 0;                       // int(p) <= 0 (also NaN)
        else if (p >= (double)scale) model = scale - 1;
        else model = (uint32_t)p;
    }
    return p;
}

// rmi_predict for N keys at once, level by level: the N chains of dependent parameter loads travel together
template <int N>
GSM_HD void rmi_predict_n(const RmiModel& m, const uint64_t* code, double* p) {
    uint32_t model[N];
    for (int j = 0; j < N; ++j) { model[j] = 0; p[j] = 0.0; }
    for (uint32_t lv = 0; lv < m.n_levels; ++lv) {
        const uint32_t scale = (lv + 1 < m.n_levels) ? m.level_size[lv + 1] : 1u;
        for (int j = 0; j < N; ++j) {
            const uint32_t k = m.level_off[lv] + model[j];
            p[j] = mul_add_nofma((double)code[j], m.coef[(size_t)k * m.stride], m.intercept[(size_t)k * m.stride]);
            if (!(p[j] >= 1.0)) model[j] = 0;
            else if (p[j] >= (double)scale) model[j] = scale - 1;
            else model[j] = (uint32_t)p[j];
        }
    }
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

// The same last-mile search as RmiTable::lookup, written as a resumable state machine with ONE probe site:
// begin() sets up the first probe, then the caller alternates  row() -> fetch {s, code64} -> feed()  until
// !pending().  A GPU thread runs many lookups back to back through the same loop, so all threads of a warp meet at
// the single fetch no matter which phase of which search each one is in (exponential gallop up / down, shared
// binary prefix, the two binary searches of RMI_LUT.py:183-184).  Probe for probe identical to RmiTable.
struct RmiSearch {
    enum : int { DONE = 0, START, UP, DOWN, SH, SH_SKIP, BIN_END, BIN_MID, BIN_SKIP };
    int state = DONE;
    bool raised = false, have_lower = false, have_upper = false, strict = false;
    int64_t n_rows = 0, n_bases = 0;
    uint32_t K = 0;
    uint64_t q = 0;
    int64_t start = 0, win = 0, ind = 0, lower = 0, upper = 0, bl = 0, bu = 0, mid = 0;
    int depth = 0;
    int64_t out_lo = 0, out_hi = -1;

    GSM_HD bool pending() const { return state != DONE; }
    GSM_HD bool hit() const { return !raised && out_hi >= out_lo; }
    // row to fetch for the pending probe (Python negative indexing already applied)
    GSM_HD uint64_t row() const { return (uint64_t)(ind < 0 ? ind + n_rows : ind); }

    GSM_HD static int64_t floor_half(int64_t sum) { return (sum >= 0) ? sum / 2 : -((-sum + 1) / 2); }

    // schedule the probe of `i` in state `st`; an index outside [-n, n) is the reference's IndexError
    GSM_HD void issue(int st, int64_t i) {
        ind = i;
        if (i < -n_rows || i >= n_rows) { raised = true; state = DONE; return; }
        state = st;
    }

    GSM_HD void begin(const RmiModel& m, uint64_t code, int64_t rows, int64_t bases) {
        n_rows = rows; n_bases = bases; K = m.K; q = code;
        raised = false; have_lower = have_upper = false; strict = false;
        out_lo = 0; out_hi = -1;
        const double pred = rmi_predict(m, code);
        if (!(pred > -9.0e18 && pred < 9.0e18)) { raised = true; state = DONE; return; }
        start = (int64_t)pred;                                  // int() truncates toward zero (RMI_LUT.py:72)
        issue(START, start);
    }

    GSM_HD void gallop_up() {                                   // RMI_LUT.py:151-163
        if (!have_upper && start + win < n_bases + 1) { const int64_t i = start + win; win *= 2; issue(UP, i); return; }
        win = 1;
        gallop_down();
    }
    GSM_HD void gallop_down() {                                 // RMI_LUT.py:166-178
        if (!have_lower && start - win >= 0) { const int64_t i = start - win; win *= 2; issue(DOWN, i); return; }
        if (!have_lower) lower = 0;
        if (!have_upper) upper = n_rows - 1;
        depth = 0;
        shared_iter();
    }
    GSM_HD void shared_iter() {                                 // common prefix of the two binary searches
        if (depth < 400 && upper - lower > 1) { mid = floor_half(lower + upper); issue(SH, mid); return; }
        begin_bin(false);
    }
    GSM_HD void begin_bin(bool st) {
        strict = st; bl = lower; bu = upper; depth = 0;
        bin_iter();
    }
    GSM_HD void bin_iter() {                                    // RMI_LUT.py:95-133, the recursion as a loop
        if (depth >= 400) { raised = true; state = DONE; return; }     // RecursionError in the reference
        if (bl == bu) { bin_done(bl); return; }
        if (bu - bl == 1) { issue(BIN_END, strict ? bu : bl); return; }
        mid = floor_half(bl + bu);
        issue(BIN_MID, mid);
    }
    GSM_HD void bin_done(int64_t r) {
        if (!strict) { out_lo = r; begin_bin(true); return; }
        out_hi = r;
        state = DONE;
    }

    // result of the pending probe: s = suffix_array[row], code64 = 32-mer code at text position s-1
    GSM_HD void feed(int64_t s, uint64_t code64) {
        const bool ok = !(s - 1 + (int64_t)K > n_bases);       // get_ref_seq returns None for a short suffix
        const uint64_t c = ok ? (code64 >> (64u - 2u * K)) : 0ull;
        switch (state) {
        case START:
            if (!ok) { start += 1; issue(START, start); return; }
            if (c < q) { lower = start; have_lower = true; }
            else if (c > q) { upper = start; have_upper = true; }
            win = 1;
            gallop_up();
            return;
        case UP:
            if (!ok) { issue(UP, ind + 1); return; }
            if (c > q) { upper = ind; have_upper = true; win = 1; gallop_down(); return; }
            if (c < q) { lower = ind; have_lower = true; }
            gallop_up();
            return;
        case DOWN:
            if (!ok) { issue(DOWN, ind - 1); return; }
            if (c < q) { lower = ind; have_lower = true; gallop_down(); return; }
            if (c > q) { upper = ind; have_upper = true; }
            gallop_down();
            return;
        case SH:
        case SH_SKIP: {
            bool at_lower = false;
            if (state == SH_SKIP && mid == lower) at_lower = true;
            else if (!ok && mid > lower) { mid -= 1; issue(SH_SKIP, mid); return; }
            if (at_lower || !ok || c == q) { begin_bin(false); return; }      // the searches part ways here
            if (c < q) lower = mid; else upper = mid;
            depth++;
            shared_iter();
            return;
        }
        case BIN_END:
            if (strict) bin_done((ok && c == q) ? bu : bl);
            else bin_done((ok && c == q) ? bl : bu);
            return;
        case BIN_MID:
        case BIN_SKIP:
            if (state == BIN_SKIP && mid == bl) {
                if (strict) { issue(BIN_END, bu); return; }
                bin_done((ok && c == q) ? bl : bu);
                return;
            }
            if (!ok && mid > bl) { mid -= 1; issue(BIN_SKIP, mid); return; }
            if (!ok) { raised = true; state = DONE; return; }                 // None < str: TypeError in the reference
            if (c < q || (c == q && strict)) bl = mid; else bu = mid;
            depth++;
            bin_iter();
            return;
        default:
            return;
        }
    }
};

// Error-bounded last-mile search for the common case: same results as the literal search above at a fraction of the
// instructions.  The literal algorithm deviates from "first row >= q, last row <= q" only when a row of its bracket is a
// None row of get_ref_seq (a suffix shorter than K: K rows of the whole table, RMI_LUT.py:89-92), when the prediction
// falls outside the table, or when a bracket end stays at its default -- its exponential phase otherwise ends with
// k-mer[lower] < q < k-mer[upper], and both binary searches of a None-free sorted range return the exact bounds whatever
// pivots they use.  So: replay the exponential phase probe for probe (same rows); declare a HAZARD (-> the caller reruns
// the window through RmiSearch) on a None row, an out-of-table start, a default bracket end or a None row anywhere
// inside the bracket (none_rows: the sorted rows of the K short suffixes, behind a 1024-region bitmap); otherwise
// finish with a lower-bound binary search and a galloping upper-bound search (k-mer counts are small).
// The search is cut into three PHASES whose per-probe update is straight-line code on 32-bit rows, so that a warp can
// run each phase as one lock-step loop with (nearly) uniform instructions: A gallop from the prediction to a bracket,
// B lower bound, C upper bound (hits only).  k_select_seeded runs phase A for all windows of a round, then B, then C;
// rmi_fast_lookup below (host test, reference for the kernel) runs them back to back for one k-mer.
struct RmiCommon {
    uint64_t q = 0;
    uint32_t K = 0, n_rows = 0;
    int64_t n_bases = 0;
    // k-mer of a probed row; false for a None row (suffix shorter than K)
    GSM_HD bool kmer(int64_t s, uint64_t code64, uint64_t& c) const {
        c = code64 >> (64u - 2u * K);
        return !(s - 1 + (int64_t)K > n_bases);
    }
};

struct RmiGallop : RmiCommon {                 // phase A: RMI_LUT.exponential_search's bracket, probe for probe
    uint32_t start = 0, win = 1, dir = 0, r = 0, lower = 0, upper = 0;
    bool busy = false, hazard = false, have_lower = false, have_upper = false;
    // int(prediction) when it is a row of the table, else -1 (the literal search wraps or raises there)
    GSM_HD static int64_t row_of(double pred, uint32_t rows) {
        if (!(pred > -1.0 && pred < (double)rows)) return -1;
        return (int64_t)pred;                                   // int() truncates toward zero (RMI_LUT.py:72)
    }
    GSM_HD static int64_t predicted_row(const RmiModel& m, uint64_t code, uint32_t rows) { return row_of(rmi_predict(m, code), rows); }
    GSM_HD void begin(const RmiModel& m, uint64_t code, int64_t row0, uint32_t rows, int64_t bases) {
        q = code; K = m.K; n_rows = rows; n_bases = bases;
        have_lower = have_upper = false; dir = 0; win = 1;
        hazard = row0 < 0;                    // prediction outside the table
        busy = !hazard;
        start = r = hazard ? 0u : (uint32_t)row0;
    }
    GSM_HD uint64_t row() const { return r; }
    GSM_HD void feed(const RmiModel& m, int64_t s, uint64_t code64) {
        uint64_t c;
        const bool ok = kmer(s, code64, c);
        const bool lt = c < q, gt = c > q;
        lower = lt ? r : lower; have_lower |= lt;
        upper = gt ? r : upper; have_upper |= gt;
        if (dir == 0u) { dir = 1u; win = 1u; }
        if (dir == 1u && (have_upper || (uint64_t)start + win >= n_rows)) { dir = 2u; win = 1u; }
        const bool done = dir == 2u && (have_lower || start < win);
        r = dir == 1u ? start + win : start - win;
        win <<= 1;
        if (!ok) { hazard = true; busy = false; return; }           // a None row: the literal search handles this window
        if (!done) return;
        busy = false;
        hazard = !have_lower || !have_upper;                         // a default bracket end
        if (hazard) return;
        const uint32_t rl = lower >> m.none_shift, ru = upper >> m.none_shift;
        if (ru - rl > 1u || ((m.none_map[rl >> 5] >> (rl & 31u)) & 1u) || ((m.none_map[ru >> 5] >> (ru & 31u)) & 1u))
            for (uint32_t t = 0; t < m.n_none; ++t) hazard |= (m.none_rows[t] >= lower && m.none_rows[t] <= upper);
    }
};

struct RmiLower : RmiCommon {                  // phase B: first row of (lower, upper] whose k-mer is >= q
    uint32_t lo = 0, hi = 0;
    bool busy = false, hi_eq = false;
    GSM_HD void begin(const RmiModel& m, uint64_t code, uint32_t lower, uint32_t upper, uint32_t rows, int64_t bases) {
        q = code; K = m.K; n_rows = rows; n_bases = bases;
        lo = lower; hi = upper; hi_eq = false;
        busy = hi - lo > 1u;
    }
    GSM_HD uint64_t row() const { return lo + ((hi - lo) >> 1); }
    GSM_HD void feed(int64_t s, uint64_t code64) {
        uint64_t c;
        kmer(s, code64, c);                    // the bracket holds no None row
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const bool lt = c < q;
        lo = lt ? mid : lo;
        hi_eq = lt ? hi_eq : (c == q);
        hi = lt ? hi : mid;
        busy = hi - lo > 1u;
    }
};

struct RmiUpper : RmiCommon {                  // phase C: k-mer[lo] == q < k-mer[hi]: last row equal to q (counts are small: gallop)
    uint32_t lo = 0, hi = 0, g = 1;
    bool busy = false;
    GSM_HD void begin(const RmiModel& m, uint64_t code, uint32_t first, uint32_t upper, uint32_t rows, int64_t bases) {
        q = code; K = m.K; n_rows = rows; n_bases = bases;
        lo = first; hi = upper; g = 1;
        busy = hi - lo > 1u;
    }
    GSM_HD uint32_t next() const { return (g != 0u && (uint64_t)lo + g < hi) ? lo + g : lo + ((hi - lo) >> 1); }
    GSM_HD uint64_t row() const { return next(); }
    GSM_HD void feed(int64_t s, uint64_t code64) {
        uint64_t c;
        kmer(s, code64, c);
        const uint32_t r = next();
        const bool eq = c == q;
        lo = eq ? r : lo;
        hi = eq ? hi : r;
        g = eq ? g << 1 : 0u;                  // first larger k-mer: switch to bisection
        busy = hi - lo > 1u;
    }
};

// The three phases back to back for one k-mer.  Returns false on a hazard (use the literal search), else (lo, hi) with
// hi >= lo <=> hit; n_probes counts the fetches.
template <typename Probe>
GSM_HD bool rmi_fast_lookup(Probe& probe, const RmiModel& m, uint64_t code, uint32_t rows, int64_t bases, int64_t& lo, int64_t& hi,
                            uint32_t* n_probes = nullptr) {
    int64_t s;
    uint64_t c64;
    uint32_t np = 0;
    RmiGallop a;
    a.begin(m, code, RmiGallop::predicted_row(m, code, rows), rows, bases);
    while (a.busy) { probe(a.row(), s, c64); a.feed(m, s, c64); ++np; }
    if (n_probes) *n_probes = np;
    if (a.hazard) return false;
    RmiLower b;
    b.begin(m, code, a.lower, a.upper, rows, bases);
    while (b.busy) { probe(b.row(), s, c64); b.feed(s, c64); ++np; }
    lo = (int64_t)b.hi;
    hi = lo - 1;                                                // absent: lo = hi + 1
    if (b.hi_eq) {
        RmiUpper c;
        c.begin(m, code, b.hi, a.upper, rows, bases);
        while (c.busy) { probe(c.row(), s, c64); c.feed(s, c64); ++np; }
        hi = (int64_t)c.lo;
    }
    if (n_probes) *n_probes = np;
    return true;
}

// Any None row of get_ref_seq (the K short suffixes) inside rows [a, b]?  Bitmap over 1024 regions first, exact list second.
GSM_HD bool rmi_none_in_range(const RmiModel& m, uint32_t a, uint32_t b) {
    const uint32_t ra = a >> m.none_shift, rb = b >> m.none_shift;
    if (rb - ra <= 1u && !((m.none_map[ra >> 5] >> (ra & 31u)) & 1u) && !((m.none_map[rb >> 5] >> (rb & 31u)) & 1u)) return false;
    bool hit = false;
    for (uint32_t t = 0; t < m.n_none; ++t) hit |= (m.none_rows[t] >= a && m.none_rows[t] <= b);
    return hit;
}

// The error-bounded search WITHOUT probes, for a k-mer whose true bounds are already known from the FM index
// (seed-table entry of its last seed_K bases + K - seed_K backward steps, continued through empty intervals so that an
// absent k-mer still yields its insertion point): A = number of rows whose suffix is smaller than q (first row >= q),
// B = A + occurrences (first row > q).  Rows are sorted, so a probed non-None row r compares as  k-mer[r] < q <=> r < A  and
// k-mer[r] > q <=> r >= B: the exponential phase of RMI_LUT.exponential_search (RMI_LUT.py:136-184) is replayed on row
// NUMBERS alone -- same probe rows as RmiGallop, no table fetch -- which yields the bracket the literal search would
// reach.  If no None row lies between the lowest and the highest probed row (they enclose the bracket) and neither
// bracket end stays at its default, both binary searches of the literal algorithm return the exact bounds (see
// RmiGallop): lo = A, hi = B - 1 (hi = lo - 1 <=> absent).  Otherwise: hazard, the caller runs the literal search.
// row0 = RmiGallop::predicted_row.  Returns false on a hazard.
// The two gallops have closed forms: going up from `start` the probes are start + 1, 2, 4, ..., the first one at or beyond
// B ends the phase (or the table does: default upper end, hazard), and a probe fell below A iff the first one did; going
// down likewise with the first power of two that exceeds start - A.
GSM_HD uint32_t clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__clz((int)x);
#else
    return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}
GSM_HD bool rmi_arith_lookup(const RmiModel& m, int64_t row0, uint32_t A, uint32_t cnt, uint32_t n_rows, int64_t& lo, int64_t& hi) {
    if (row0 < 0) return false;                                   // prediction outside the table
    const uint32_t start = (uint32_t)row0, B = A + cnt;
    bool have_lower = start < A;
    uint32_t rmin = start, rmax = start;
    if (start < B) {                                              // RMI_LUT.py:151-163: first w = 2^t with start + w >= B
        const uint32_t d = B - start;
        const uint64_t w = 1ull << (d == 1u ? 0u : 32u - clz32(d - 1u));
        if (w >= (uint64_t)n_rows - start) return false;          // the table ends first: the upper end stays at its default
        rmax = start + (uint32_t)w;
        have_lower |= start + 1u < A;                             // the first (lowest) probe of the phase
    }
    if (!have_lower) {                                            // RMI_LUT.py:166-178: first w = 2^t with start - w < A
        const uint32_t g = start - A;
        const uint64_t w = 1ull << (g == 0u ? 0u : 32u - clz32(g));
        if (w > start) return false;                              // row 0 comes first: the lower end stays at its default
        rmin = start - (uint32_t)w;
    }
    if (rmi_none_in_range(m, rmin, rmax)) return false;           // a None row probed or inside the bracket
    lo = (int64_t)A;
    hi = (int64_t)B - 1;
    return true;
}

// the same replay probe by probe (the loops of RMI_LUT.exponential_search on row numbers): what the closed form is tested against
GSM_HD bool rmi_arith_lookup_loops(const RmiModel& m, int64_t row0, uint32_t A, uint32_t cnt, uint32_t n_rows, int64_t& lo, int64_t& hi) {
    if (row0 < 0) return false;
    const uint32_t start = (uint32_t)row0, B = A + cnt;
    bool have_lower = start < A, have_upper = start >= B;
    uint32_t rmin = start, rmax = start;
    for (uint32_t w = 1; !have_upper && (uint64_t)start + w < n_rows && w != 0u; w <<= 1) {
        rmax = start + w;
        have_upper = rmax >= B;
        have_lower |= rmax < A;
    }
    for (uint32_t w = 1; !have_lower && start >= w && w != 0u; w <<= 1) {
        rmin = start - w;
        have_lower = rmin < A;
        have_upper |= rmin >= B;
    }
    if (!have_lower || !have_upper) return false;
    if (rmi_none_in_range(m, rmin, rmax)) return false;
    lo = (int64_t)A;
    hi = (int64_t)B - 1;
    return true;
}

// ------------------------------------------------------------------------------------ hazard codes of a model
// The result of RMI_LUT.get_suffix_rmi for a k-mer (RMI_LUT.py:67-78) is a function of the k-mer's CODE alone: the model's
// prediction, the k-mer's true bounds and the K None rows are all fixed per (model, reference).  A code is a HAZARD iff
// rmi_arith_lookup cannot certify that the literal search returns the true interval; every other code looks up exactly.
// On a read none of whose windows is a hazard, get_smems_rmi (SMEM.py:206-384) sees at every lookup what get_smems_lut
// (SMEM.py:20-192) with the same K sees -- hit iff the k-mer occurs, its true interval, and check_sequential over the same
// positions (get_positions of the true rows = the table's position list) -- and the two routines are the same text line for
// line apart from the lookup; get_smems_lut in turn emits get_SMEMS's records with min_len 1.  Such a
// read therefore takes its RMI-SMEM records from the BWA-SMEM selection, and only reads with a hazard window run the frame
// machine.  The hazard codes (a few thousand of 4^15 for a model worth using) are kept in an open-addressing hash set of
// 32-bit codes (K <= 15: a code never equals HZ_EMPTY), at most half full, so a probe sequence always ends at an empty slot.
constexpr uint32_t HZ_EMPTY = 0xFFFFFFFFu;

GSM_HD bool rmi_code_is_hazard(const RmiModel& m, uint64_t code, uint32_t A, uint32_t cnt, uint32_t n_rows) {
    int64_t lo, hi;
    return !rmi_arith_lookup(m, RmiGallop::predicted_row(m, code, n_rows), A, cnt, n_rows, lo, hi);
}

GSM_HD uint32_t hz_slot(uint32_t code, uint32_t mask) { return ((code * 0x9E3779B1u) >> 8) & mask; }

// slot(h) = h-th word of the table; mask = slots - 1 (a power of two)
template <typename LoadSlot>
GSM_HD bool hz_contains(LoadSlot slot, uint32_t mask, uint32_t code) {
    for (uint32_t h = hz_slot(code, mask);; h = (h + 1u) & mask) {
        const uint32_t v = slot(h);
        if (v == code) return true;
        if (v == HZ_EMPTY) return false;
    }
}

// host: fill `slots` (n_slots words, a power of two >= 2 n + 2, at most 2^24) with the n codes; false if the table cannot take them
inline bool hz_build(const uint32_t* codes, uint64_t n, uint32_t* slots, uint32_t n_slots) {
    if (n_slots < 2u || (n_slots & (n_slots - 1u)) != 0u || n_slots > (1u << 24) || 2u * n + 2u > (uint64_t)n_slots) return false;
    const uint32_t mask = n_slots - 1u;
    for (uint32_t h = 0; h < n_slots; ++h) slots[h] = HZ_EMPTY;
    for (uint64_t k = 0; k < n; ++k) {
        if (codes[k] == HZ_EMPTY) return false;
        uint32_t h = hz_slot(codes[k], mask);
        while (slots[h] != HZ_EMPTY && slots[h] != codes[k]) h = (h + 1u) & mask;
        slots[h] = codes[k];
    }
    return true;
}

// No window q[i:i+K), 0 <= i <= L-K, of the read is a hazard code (a read shorter than K has no window).  rd(w) = packed read
// word w (2 bits per base, MSB first); the code of the window ending at base i is rolled from the one ending at i-1.  K <= 15.
template <typename ReadW, typename LoadSlot>
GSM_HD bool rmi_read_hazard_free(ReadW rd, uint32_t L, uint32_t K, LoadSlot slot, uint32_t mask) {
    const uint32_t kmask = (1u << (2u * K)) - 1u;
    uint32_t code = 0, w = 0;
    for (uint32_t i = 0; i < L; ++i) {
        if ((i & 15u) == 0u) w = rd(i >> 4);
        code = ((code << 2) | ((w >> (30u - 2u * (i & 15u))) & 3u)) & kmask;
        if (i + 1u >= K && hz_contains(slot, mask, code)) return false;
    }
    return true;
}

// True bounds (A, occurrences) of the K-mer q[cpos:cpos+K) from the sweep's seed table: entry of its last seed_K bases,
// then K - seed_K backward steps.  The table stores the insertion point in fwd_lo for absent k-mers (cnt == 0) and a
// backward step on an empty interval keeps tracking it, so (A, 0) is exact for absent k-mers too.
// rd(w) = packed read word w; load(i) = i-th 32-byte bucket half; seed(code) = table entry.  Needs seed_K <= K.
template <typename ReadW, typename LoadHalf, typename LoadSeed>
GSM_HD void kmer_bounds_seeded(ReadW rd, LoadHalf load, LoadSeed seed, const IndexMeta& meta, uint32_t cpos, uint32_t K, uint32_t seed_K,
                               uint32_t& A, uint32_t& cnt) {
    const U4 e = seed(kmer_code(rd, (uint64_t)cpos + K - seed_K, seed_K));
    A = e.x; cnt = e.y;
    for (uint32_t p = cpos + K - seed_K; p > cpos; --p) {
        const uint32_t w = rd((p - 1) >> 4);
        const uint32_t c = (w >> (30u - 2u * ((p - 1) & 15u))) & 3u;
        const StepOut r = step_single(load, A, A + cnt, c, meta.C[c], meta.prim_f);
        A = r.lo_new; cnt = r.cnt_new;
    }
}

// check_sequential (SMEM.py:196-202) on the positions of two RETURNED intervals [clo, chi] and [plo, phi] (SMEM.py:262-265),
// which may be wrong: is there x in the first and y in the second with suffix_array[x] + 1 == suffix_array[y]?
// suffix_array[x] + 1 == suffix_array[y]  <=>  x == LF(y), and LF maps the rows of [plo, phi] whose BWT symbol is b onto the
// contiguous rows C[b] + rank(b, plo) .. C[b] + rank(b, phi + 1) - 1: four rank pairs on the same two buckets, O(1) whatever
// the interval sizes (a poly-A k-mer on a human-scale text has 10^5..10^6 rows; the literal double loop would be 10^10+
// dependent suffix-array reads in one thread).  The '$' row has no predecessor and is excluded by the rank correction.
// Rows outside [0, n) (Python negative indexing of a wrong interval) take the literal loop, as the reference does.
template <typename LoadHalf, typename LoadSa>
GSM_HD bool rmi_sequential(LoadHalf load, LoadSa sa, const IndexMeta& meta, int64_t clo, int64_t chi, int64_t plo, int64_t phi) {
    const int64_t n = (int64_t)meta.n_rows;
    if (clo >= 0 && plo >= 0 && chi < n && phi < n) {
        if (chi < clo || phi < plo) return false;
        for (uint32_t b = 0; b < 4u; ++b) {
            const StepOut r = step_single(load, (uint32_t)plo, (uint32_t)phi + 1u, b, meta.C[b], meta.prim_f);
            if (r.cnt_new != 0u && (int64_t)r.lo_new <= chi && (int64_t)r.lo_new + r.cnt_new - 1 >= clo) return true;
        }
        return false;
    }
    for (int64_t x = clo; x <= chi; ++x) {
        const uint32_t px = sa((uint64_t)(x < 0 ? x + n : x));
        for (int64_t y = plo; y <= phi; ++y)
            if (px + 1u == sa((uint64_t)(y < 0 ? y + n : y))) return true;
    }
    return false;
}

// check_sequential on two TRUE k-mer intervals (LUT entries, or RMI lookups proven exact): some occurrence of q[c:c+K) is
// followed one base later by q[pc:pc+K)  <=>  the k-mers overlap consistently and q[c] + q[pc:pc+K) occurs (one backward
// step on the second seed).
template <typename Base, typename LoadHalf>
GSM_HD bool true_sequential(Base base, LoadHalf load, const IndexMeta& meta, uint32_t K, uint32_t c, uint32_t pc, int64_t plo, int64_t phi) {
    for (uint32_t t = 1; t < K; ++t)
        if (base(c + t) != base(pc + t - 1)) return false;
    const uint32_t ch = base(c);
    const StepOut r = step_single(load, (uint32_t)plo, (uint32_t)phi + 1u, ch, meta.C[ch], meta.prim_f);
    return r.cnt_new != 0;
}

// ------------------------------------------------------------------------------------ selection
// Ctx must provide:
//   uint32_t L, K, n_mems;  uint32_t min_len;
//   MemEntry mem(uint32_t k)                 k-th maximal match, sorted by end (and start)
//   uint32_t se(uint32_t k)                  its start | end << 16 word alone
//   uint32_t base(uint32_t pos)
//   uint32_t seed_round(bool first, uint32_t e, uint32_t plen, uint32_t nwin, W& win, uint32_t& wtrue)
//                                            (only for Selector::run_seeded) LUT / RMI lookups of the windows of one
//                                            round: window i covers q[c:c+K), c = first ? 0 : e - i, and is visited iff
//                                            first or (i < plen and c + K <= L); win.put(i, lo, hi) stores its result;
//                                            bit i of the return value = hit; bit i of wtrue = the stored pair is the
//                                            k-mer's TRUE interval (LUT: always)
//   bool sequential(uint32_t c, int64_t clo, int64_t chi, uint32_t pc, int64_t plo, int64_t phi, bool both_true)
//                                            check_sequential of the two seeds (SMEM.py:196-202)
//   void interval(uint32_t i, uint32_t j, uint32_t& lo, uint32_t& cnt)   true SA interval of q[i:j]
//   void emit(uint32_t i, uint32_t j, int64_t lo, int64_t hi)
//   typename iv_t                             type of SA rows in seed tuples (uint32_t, or int64_t if rows may be negative)
//   bool failed()                             the reference raised inside seed()
//   bool seeds_are_true()                     seed() returns exact SA intervals (LUT), not RMI guesses
constexpr int MAX_SEED_K = 32;      // LUT K <= 16, RMI K <= 26 (float64-exact codes)

template <typename Ctx>
struct Selector {
    // SA rows of a seed / candidate: Ctx::iv_t is uint32_t where rows are always real (LUT) and int64_t where the RMI search
    // may return negative or wrapped rows
    using iv_t = typename Ctx::iv_t;
    GSM_HD static uint32_t s_of(const MemEntry& e) { return e.se & 0xFFFFu; }
    GSM_HD static uint32_t e_of(const MemEntry& e) { return e.se >> 16; }
    // start / end of match k from its packed (start | end << 16) word alone: Ctx::se(k) may come from a compact per-read
    // copy (the lookups below only need these 4 bytes of the 16-byte entries)
    GSM_HD static uint32_t sk(Ctx& c, uint32_t k) { return c.se(k) & 0xFFFFu; }
    GSM_HD static uint32_t ek(Ctx& c, uint32_t k) { return c.se(k) >> 16; }

    // Starts and ends of the maximal matches both increase strictly with k, so every lookup into the list is a
    // binary search: first k with e_k > p (or >= with `incl`), and number of k with s_k <= p.
    GSM_HD static uint32_t first_end_above(Ctx& c, uint32_t p, uint32_t lo = 0) {
        uint32_t hi = c.n_mems;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ek(c, mid) <= p) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
    GSM_HD static uint32_t count_starts_upto(Ctx& c, uint32_t p) {
        uint32_t lo = 0, hi = c.n_mems;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sk(c, mid) <= p) lo = mid + 1; else hi = mid;
        }
        return lo;
    }

    // get_SMEM_at_index (SMEM.py:469-484) == longest maximal match covering p, ties to the
    // smallest end (the strict '>' of SMEM.py:413 scanning ends upward).  `from` = first k with
    // e_k > p.  Returns the index of the winner.
    GSM_HD static uint32_t covering_best(Ctx& c, uint32_t p, uint32_t& from) {
        from = first_end_above(c, p, from);
        uint32_t best = from, bestlen = 0;
        for (uint32_t k = from; k < c.n_mems; ++k) {
            const uint32_t w = c.se(k), st = w & 0xFFFFu;
            if (st > p) break;
            const uint32_t len = (w >> 16) - st;
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

    // A candidate record of the frame machine: the match q[i:j) and WHERE its SA interval can be read.  Only LENGTHS decide
    // which candidate survives (SMEM.py replaces on >=), so a candidate is two words -- (i | j << 16) and the source of its
    // interval: a window of the round (the seed tuple itself), an entry of the match list, or "lazy" = neither, computed once
    // for the round's winner only (resolve): at most one explicit backward search per record.
    struct Cand {
        uint32_t ij = 0, src = 0;                       // ij == 0: no candidate (a candidate has j > 0)
        GSM_HD bool valid() const { return ij != 0u; }
        GSM_HD uint32_t i() const { return ij & 0xFFFFu; }
        GSM_HD uint32_t j() const { return ij >> 16; }
        GSM_HD uint32_t len() const { return j() - i(); }
    };
    static constexpr uint32_t SRC_LAZY = 0xFFFFFFFFu, SRC_WIN = 0x80000000u;     // else: index into the match list

    GSM_HD static void upd(Cand& cd, const Cand& x) {
        if (!cd.valid() || x.len() >= cd.len()) cd = x;
    }
    GSM_HD static Cand seed_key(uint32_t i, uint32_t K, uint32_t w) { return Cand{i | ((i + K) << 16), SRC_WIN | w}; }
    GSM_HD static Cand listed(uint32_t i, uint32_t j, uint32_t k) { return Cand{i | (j << 16), k}; }
    GSM_HD static Cand lazy_iv(uint32_t i, uint32_t j) { return Cand{i | (j << 16), SRC_LAZY}; }

    // F(p) restricted to true matches = end of the last match starting at or before p; 0 such matches => p
    // (cannot happen when all four bases occur)
    GSM_HD static uint32_t F_of(Ctx& c, uint32_t p) {
        const uint32_t n = count_starts_upto(c, p);
        if (n == 0) return p;
        const uint32_t e = ek(c, n - 1);
        return e > p ? e : p;
    }

    // True SA interval of q[i:j): read it off the match list when (i, j) is itself a maximal match
    // (the usual case); false = one backward search from j down to i is needed (Ctx::interval).
    GSM_HD static bool listed_iv(Ctx& c, uint32_t i, uint32_t j, iv_t& lo, iv_t& hi) {
        const uint32_t k = j ? first_end_above(c, j - 1) : 0u;
        if (k < c.n_mems && c.se(k) == (i | (j << 16))) {
            const MemEntry m = c.mem(k);
            lo = (iv_t)m.lo; hi = (iv_t)(m.lo + m.cnt - 1u);
            return true;
        }
        return false;
    }

    // winner of a round -> its interval, from the round's window results W (lo(i) / hi(i)), the match list, or (returns
    // true) one explicit backward search from j down to i (Ctx::interval) that the caller still has to run.
    template <typename W>
    GSM_HD static bool resolve(Ctx& c, const Cand& cd, const W& win, iv_t& lo, iv_t& hi) {
        if (cd.src == SRC_LAZY) return !listed_iv(c, cd.i(), cd.j(), lo, hi);
        if (cd.src & SRC_WIN) { lo = win.lo(cd.src & 0xFFu); hi = win.hi(cd.src & 0xFFu); return false; }
        const MemEntry m = c.mem(cd.src);
        lo = (iv_t)m.lo; hi = (iv_t)(m.lo + m.cnt - 1u);
        return false;
    }

    // forward_extension(query, pc+K, kmer, seed) (SMEM.py:425-443): longest key and its value.  f = F(pc), w = the
    // window whose lookup is the frame's seed tuple.
    GSM_HD static Cand fwd_only(Ctx& c, uint32_t pc, uint32_t w, uint32_t f) {
        if (f <= pc + c.K) return seed_key(pc, c.K, w);              // the seed key itself
        return lazy_iv(pc, f);
    }

    // backward_extension(query, pc, keys) (SMEM.py:389-423) over keys pc+K .. max(F(pc), pc+K)
    // (all_keys) or over the seed key only.  f = F(pc), k0 = first match whose end exceeds pc + K - 1.
    GSM_HD static Cand bext(Ctx& c, uint32_t pc, uint32_t w, bool all_keys, uint32_t f, uint32_t k0) {
        const uint32_t K = c.K;
        const bool seed_true = f >= pc + K;           // the k-mer really occurs
        uint32_t jmax = all_keys ? (f > pc + K ? f : pc + K) : pc + K;
        // extended keys: j in [pc+K, jmax] with LS[j] < pc; only maximal-match ends and jmax matter
        bool have = false;
        uint32_t bi = 0, bj = 0, bk = 0;
        bool b_from_mem = false;
        if (seed_true) {
            for (uint32_t k = k0; k < c.n_mems; ++k) {
                const uint32_t v = c.se(k), s = v & 0xFFFFu, e = v >> 16;
                if (s >= pc) break;                    // starts are sorted: no further left extension
                uint32_t j = e <= jmax ? e : jmax;     // plateau end, or the key range's last key
                if (!have || (j - s) > (bj - bi)) {
                    have = true; bi = s; bj = j; bk = k;
                    b_from_mem = (j == e);
                }
                if (e >= jmax) break;
            }
        }
        // the longest key wins only if strictly longer (SMEM.py:418)
        uint32_t fend = jmax;
        if (!have || (fend - pc) > (bj - bi)) {
            if (fend == pc + K) return seed_key(pc, K, w);
            return lazy_iv(pc, fend);
        }
        if (!b_from_mem) return lazy_iv(bi, bj);
        return listed(bi, bj, bk);
    }

    // get_smems_lut / get_smems_rmi: the frame machine of SMEM.py:49-186 / :235-379, one ROUND at a time.  A round =
    // the windows left of the previous SMEM's end (the first round: window 0 only) and emits exactly one record.
    //   Pass 1 (Ctx::seed_round): the lookups of ALL windows of the round.  The reference visits every window whatever
    //   the earlier ones returned (no early exit in SMEM.py:56-146), so the lookups can be taken out of the machine and
    //   run where the threads of a warp are together (k_select_seeded keeps a warp in lock step on them).
    //   Pass 2 (round_decide): the machine over the stored results -> the round's winner, its interval possibly lazy.
    //   Pass 3 (resolve + Ctx::interval, at most one explicit backward search) and round_commit (emit, advance).
    struct Seeded {
        bool first = true, done = false;
        uint32_t e = 0, plen = 0;
    };

    // false once the read is finished (SMEM.py:49 `while e < len(query)`)
    GSM_HD static bool round_needed(const Ctx& c, Seeded& st) {
        if (st.done || (!st.first && st.e >= c.L)) { st.done = true; return false; }
        return true;
    }

    // the lookup results of one round, per window: plain arrays here; k_select_seeded keeps them in shared memory
    struct Windows {
        iv_t l[MAX_SEED_K], h[MAX_SEED_K];
        GSM_HD iv_t lo(uint32_t i) const { return l[i]; }
        GSM_HD iv_t hi(uint32_t i) const { return h[i]; }
        GSM_HD void put(uint32_t i, iv_t lo_, iv_t hi_) { l[i] = lo_; h[i] = hi_; }
    };

    GSM_HD static void run_seeded(Ctx& c) {
        Seeded st;
        Windows win;
        while (round_needed(c, st)) {
            uint32_t wtrue = 0;
            const uint32_t whit = c.seed_round(st.first, st.e, st.plen, st.first ? 1u : c.K, win, wtrue);
            if (c.failed()) return;
            const Cand w = round_decide(c, st, win, whit, wtrue);
            iv_t lo = 0, hi = 0;
            if (w.valid() && resolve(c, w, win, lo, hi)) {
                uint32_t l, n;
                c.interval(w.i(), w.j(), l, n);
                lo = (iv_t)l; hi = (iv_t)(l + n - 1u);
            }
            round_commit(c, st, w, lo, hi);
        }
    }

    // emit the round's winner and advance; an invalid winner ends the read (no match covers e: only when a base of the read
    // is absent from the text, outside the reference's domain)
    GSM_HD static void round_commit(Ctx& c, Seeded& st, const Cand& w, iv_t lo, iv_t hi) {
        if (!w.valid()) { st.done = true; return; }
        c.emit(w.i(), w.j(), lo, hi);
        st.first = false;
        st.e = w.j(); st.plen = w.len();
    }

    // wtrue: bit i set = window i's (lo, hi) is the TRUE interval of its k-mer (always for LUT; for RMI when the lookup
    // was proven exact), which lets check_sequential be evaluated in closed form.
    template <typename W>
    GSM_HD static Cand round_decide(Ctx& c, const Seeded& st, const W& win, uint32_t whit, uint32_t wtrue) {
        const uint32_t K = c.K, L = c.L;
        const uint32_t e = st.e, plen = st.plen;
        if (st.first) {                                    // SMEM.py:26-39 / :213-225
            if (whit & 1u) return fwd_only(c, 0, 0, F_of(c, 0));
            const uint32_t end = F_of(c, 0);
            if (end == 0) return Cand{};
            return lazy_iv(0, end);
        }
        // frame: 0 = None, 1 = () , 2 = k-mer frame (start pc, seed tuple = the lookup of window pw)
        int fstate = 0;
        uint32_t pc = 0, pw = 0; bool pfw = false;
        uint32_t pF = 0, pE = 0;                 // of the frame's window: F(pc) and the first match whose end exceeds pc + K - 1
        Cand cd{};
        const uint32_t pstart = e - plen;
        // The windows of a round move left one base at a time, so the two positions the machine looks up in the match list
        // per window (starts <= cpos, ends > cpos + K - 1) are CURSORS that only step left: two binary searches per round
        // instead of three per window.
        uint32_t nf = count_starts_upto(c, e), ne = first_end_above(c, e + K - 1);
        // Pass 2: the frame machine over the stored results.  Each window only DECIDES what happens to the previous frame
        // (A_*); the extension itself runs at one place below, so that the threads of a warp that extend a frame do it
        // together whichever case of the reference they are in.  Step i == K is the closing extension (SMEM.py:149-171).
        enum { A_NONE = 0, A_BEXT = 1, A_FWD = 2, A_KNOWN = 3 };
        for (uint32_t i = 0; i <= K; ++i) {
            int act = A_NONE;
            uint32_t ai = 0, aw = 0;                // the frame (start, window of its seed) the action applies to
            uint32_t aF = 0, aE = 0;                // A_BEXT / A_FWD: the frame's F(pc) and first-end cursor
            bool aall = false;
            if (i == K) {
                if (fstate == 2) { act = A_BEXT; ai = pc; aw = pw; aall = pfw; aF = pF; aE = pE; }
            } else {
                if (i >= plen) continue;
                const uint32_t cpos = e - i;
                if (cpos + K > L) continue;
                while (nf > 0u && sk(c, nf - 1u) > cpos) --nf;                    // matches starting at or before cpos
                while (ne > 0u && ek(c, ne - 1u) > cpos + K - 1u) --ne;           // first match ending beyond the window
                uint32_t Fc = cpos;                                               // F(cpos): end of the longest match starting there
                if (nf > 0u) { const uint32_t en = ek(c, nf - 1u); Fc = en > cpos ? en : cpos; }
                const bool hit = (whit >> i) & 1u;
                if (hit) {
                    if (fstate == 0) { fstate = 2; pc = cpos; pw = i; pfw = true; pF = Fc; pE = ne; }          // :70
                    else if (fstate == 1) { fstate = 2; pc = cpos; pw = i; pfw = false; pF = Fc; pE = ne; }    // :73
                    else {
                        // check_sequential of two TRUE k-mer intervals at adjacent windows is just
                        // "q[cpos : cpos+K+1) occurs" (SURVEY A13): read it off the match list
                        const bool both = ((wtrue >> i) & (wtrue >> pw) & 1u) != 0u;
                        const bool seq = (both && pc == cpos + 1) ? (Fc >= cpos + K + 1)
                                                                  : c.sequential(cpos, (int64_t)win.lo(i), (int64_t)win.hi(i), pc,
                                                                                 (int64_t)win.lo(pw), (int64_t)win.hi(pw), both);
                        if (seq) {                                                                        // Case 1
                            if (!pfw && cd.valid() && (pc - pstart) + K < cd.len()) continue;             // :94-95 (the frame stays)
                            act = A_BEXT; ai = pc; aw = pw; aall = pfw; aF = pF; aE = pE;
                        } else if (pfw) {                                                                 // Case 2
                            act = A_FWD; ai = pc; aw = pw; aF = pF;
                        } else {
                            act = A_KNOWN; ai = cpos; aw = i;
                        }
                        pc = cpos; pw = i; pfw = false; pF = Fc; pE = ne;
                    }
                } else {
                    if (fstate == 2) {                                                                    // Case 3
                        if (pfw) { act = A_FWD; ai = pc; aw = pw; aF = pF; }
                        else { act = A_KNOWN; ai = pc; aw = pw; }
                    }
                    fstate = 1;
                }
            }
            if (act == A_BEXT) upd(cd, bext(c, ai, aw, aall, aF, aE));
            else if (act == A_FWD) upd(cd, fwd_only(c, ai, aw, aF));
            else if (act == A_KNOWN) upd(cd, seed_key(ai, K, aw));
        }
        if (!cd.valid()) {                                                                            // :175-179
            uint32_t from = 0;
            const uint32_t b = covering_best(c, e, from);
            if (b >= c.n_mems) return cd;
            const uint32_t v = c.se(b);
            return listed(v & 0xFFFFu, v >> 16, b);
        }
        return cd;
    }
};

}  // namespace gsm
