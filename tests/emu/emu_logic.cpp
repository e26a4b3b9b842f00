// TEST INFRASTRUCTURE -- host build of the kernels' per-read logic (sweep_logic.cuh,
// select_logic.cuh, fm_core.cuh) so that the control flow can be checked against the oracle in
// the GPU-less build container.  It is compiled by tests/emu/build.py into tests/emu/_build/ and
// loaded only by tests/test_logic_emu.py; the product (genie_smem_b200) never links or calls it.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../genie_smem_b200/csrc/select_logic.cuh"

using namespace gsm;

namespace {

uint64_t g_cnt[8];   // [0] seed lookups [1] interval() calls [2] interval() FM steps [3] sequential() calls [4] reads

struct HostIndex {
    const Half* fwd;
    const Half* rev;
    const uint32_t* sa;
    const uint32_t* text;
    IndexMeta meta;
    uint64_t n_bases;
};

struct SweepCtx {
    const uint32_t* words = nullptr;
    uint32_t K = 0;                      // seed table K (0 = none)
    bool use_uniq = false;               // unique-match shortcut, forward (suffix array + text)
    bool use_uniq_back = false;          // ... and to the left (inverse suffix array)
    bool uniq() const { return use_uniq; }
    bool uniq_back() const { return use_uniq_back; }
    uint32_t seed_k() const { return K; }
    uint32_t kmer(uint32_t pos) const {
        auto rd = [&](uint64_t w) { return words[w]; };
        return (uint32_t)kmer_code(rd, pos, K);
    }
    std::vector<uint32_t> cj, clo, ccnt;
    std::vector<MemEntry> mems;
    bool fetched = false;
    uint32_t L = 0;
    bool fetch(uint32_t& rid, uint32_t& len) {
        if (fetched) return false;
        fetched = true; rid = 0; len = L;
        return true;
    }
    uint32_t base(uint32_t pos) const { return base_msb(words, pos); }
    void cand_put(uint32_t i, uint32_t j, uint32_t lo, uint32_t cnt) {
        if (i >= cj.size()) { cj.resize(i + 1); clo.resize(i + 1); ccnt.resize(i + 1); }
        cj[i] = j; clo[i] = lo; ccnt[i] = cnt;
    }
    void cand_get(uint32_t i, uint32_t& j, uint32_t& lo, uint32_t& cnt) const { j = cj[i]; lo = clo[i]; cnt = ccnt[i]; }
    void cand_sync() {}
    void emit(uint32_t idx, MemEntry e) { if (idx >= mems.size()) mems.resize(idx + 1); mems[idx] = e; }
    void finish(uint32_t, uint32_t n) { mems.resize(n); }
};

void sort_mems(std::vector<MemEntry>& m) {
    // within a sweep the ends are emitted in descending order; sweeps ascend
    size_t a = 0;
    while (a < m.size()) {
        size_t b = a;
        while (b < m.size() && m[b].sweep == m[a].sweep) ++b;
        std::reverse(m.begin() + a, m.begin() + b);
        a = b;
    }
}

struct SelCtx {
    using iv_t = int64_t;
    const HostIndex* ix;
    const uint32_t* words;
    uint32_t L, K, n_mems, min_len;
    const MemEntry* mems;
    int method;
    const uint32_t* lut;
    RmiModel rmi;
    bool raised = false;
    uint32_t seed_K = 0;                 // sweep seed table (optional): lets RMI lookups use rmi_arith_lookup
    const U4* seed_tab = nullptr;
    std::vector<uint32_t>* out;   // 6 x u32 per record: i, j, lo_lo, lo_hi, hi_lo, hi_hi

    MemEntry mem(uint32_t k) const { return mems[k]; }
    uint32_t se(uint32_t k) const { return mems[k].se; }
    uint32_t base(uint32_t pos) const { return base_msb(words, pos); }
    bool failed() const { return raised; }
    bool seeds_are_true() const { return method == GSM_METHOD_LUT_; }

    void interval(uint32_t i, uint32_t j, uint32_t& lo, uint32_t& cnt) {
        g_cnt[1]++; g_cnt[2] += j - i;
        lo = 0; cnt = ix->meta.n_rows;
        auto load = [&](uint64_t idx) { return ix->fwd[idx]; };
        uint32_t p = j;
        if (method == GSM_METHOD_LUT_ && j - i >= K) {
            auto rd = [&](uint64_t w) { return words[w]; };
            uint64_t code = kmer_code(rd, j - K, K);
            lo = lut[2 * code]; cnt = lut[2 * code + 1]; p = j - K;
            if (cnt == 0) return;
        }
        for (; p > i; --p) {
            uint32_t c = base(p - 1);
            StepOut r = step_single(load, lo, lo + cnt, c, ix->meta.C[c], ix->meta.prim_f);
            lo = r.lo_new; cnt = r.cnt_new;
            if (cnt == 0) return;
        }
    }

    template <typename W>
    uint32_t seed_round(bool first, uint32_t e, uint32_t plen, uint32_t nwin, W& win, uint32_t& wtrue) {
        auto rd = [&](uint64_t w) { return words[w]; };
        auto load = [&](uint64_t idx) { return ix->fwd[idx]; };
        auto seed = [&](uint64_t code) { return seed_tab[code]; };
        wtrue = 0;
        auto sa = [&](uint64_t r) { return ix->sa[r]; };
        auto tx = [&](uint64_t w) { return ix->text[w]; };
        SaTextProbe<decltype(sa), decltype(tx)> pr{sa, tx};
        uint32_t hit = 0;
        for (uint32_t i = 0; i < nwin; ++i) {
            const uint32_t cpos = first ? 0u : e - i;
            if (!(first || (i < plen && cpos + K <= L))) continue;
            g_cnt[0]++;
            uint64_t code = kmer_code(rd, cpos, K);
            if (method == GSM_METHOD_LUT_) {
                uint32_t l = lut[2 * code], n = lut[2 * code + 1];
                win.put(i, (int64_t)l, (int64_t)l + n - 1);
                wtrue |= 1u << i;
                if (n != 0) hit |= 1u << i;
                continue;
            }
            if (rmi.n_none) {                            // error-bounded fast path first, literal search on a hazard
                int64_t flo, fhi;
                bool ok;
                if (seed_tab && seed_K <= K) {           // no probes: true bounds from the seed table + arithmetic replay
                    uint32_t A, n;
                    kmer_bounds_seeded(rd, load, seed, ix->meta, cpos, K, seed_K, A, n);
                    ok = rmi_arith_lookup(rmi, RmiGallop::predicted_row(rmi, code, ix->meta.n_rows), A, n, ix->meta.n_rows, flo, fhi);
                    g_cnt[6]++;                          // [6] arithmetic lookups (select runs)
                } else {
                    ok = rmi_fast_lookup(pr, rmi, code, ix->meta.n_rows, (int64_t)ix->n_bases, flo, fhi);
                }
                if (ok) {
                    win.put(i, flo, fhi);
                    wtrue |= 1u << i;
                    if (fhi >= flo) hit |= 1u << i;
                    continue;
                }
                g_cnt[5]++;                              // [5] hazards (select runs)
            }
            RmiSearch rs;
            rs.begin(rmi, code, (int64_t)ix->meta.n_rows, (int64_t)ix->n_bases);
            while (rs.pending()) {
                int64_t s; uint64_t c64;
                pr(rs.row(), s, c64);
                rs.feed(s, c64);
            }
            if (rs.raised) { raised = true; return hit; }
            win.put(i, rs.out_lo, rs.out_hi);
            if (rs.hit()) hit |= 1u << i;
        }
        return hit;
    }

    bool sequential(uint32_t c, int64_t clo, int64_t chi, uint32_t pc, int64_t plo, int64_t phi, bool both_true) {
        g_cnt[3]++;
        auto load = [&](uint64_t idx) { return ix->fwd[idx]; };
        auto bs = [&](uint32_t p) { return base(p); };
        if (method == GSM_METHOD_LUT_ || both_true) return true_sequential(bs, load, ix->meta, K, c, pc, plo, phi);
        auto sal = [&](uint64_t r) { return ix->sa[r]; };
        return rmi_sequential(load, sal, ix->meta, clo, chi, plo, phi);
    }

    void emit(uint32_t i, uint32_t j, int64_t lo, int64_t hi) {
        out->push_back(i); out->push_back(j);
        out->push_back((uint32_t)lo); out->push_back((uint32_t)((uint64_t)lo >> 32));
        out->push_back((uint32_t)hi); out->push_back((uint32_t)((uint64_t)hi >> 32));
    }

    static constexpr int GSM_METHOD_LUT_ = 1;
};

}  // namespace

extern "C" {
struct EmuIndex {
    const void* fwd; const void* rev; const uint32_t* sa; const uint32_t* text;
    uint32_t C[4]; uint32_t cnt[4]; uint32_t prim_f, prim_r, n_rows, uniq_mode; uint64_t n_bases;     // uniq_mode: 0 off, 1 forward only, 2 forward + backward
    const uint32_t* isa;                 // inverse suffix array (uniq_mode 2)
};
}

// the non-FM operations of the sweep's unique-match shortcut; true if one was served
template <typename Sw, typename Ctx>
static bool serve_uniq(Sw& sw, Ctx& ctx, const IndexMeta& meta, const EmuIndex* ei, const uint32_t* words) {
    if (sw.pending_word()) {
        g_cnt[3]++;                                                          // [3] SA / ISA fetches
        sw.consume_word(ctx, meta, sw.mode == M_ISA ? ei->isa[sw.aux] : ei->sa[sw.aux]);
        return true;
    }
    if (sw.pending_cmp()) {
        g_cnt[2]++;                                                          // [2] text comparisons
        uint32_t mx = sw.cmp_max((uint32_t)ei->n_bases);
        if (mx > SWEEP_CMP_CHUNK) mx = SWEEP_CMP_CHUNK;
        const uint32_t t = sw.cmp_text(), q = sw.cmp_read();
        uint32_t mt = 0;
        if (sw.mode == M_CMPF) while (mt < mx && base_msb(ei->text, t + mt) == base_msb(words, q + mt)) ++mt;
        else while (mt < mx && base_msb(ei->text, t - 1 - mt) == base_msb(words, q - 1 - mt)) ++mt;
        sw.consume_cmp(ctx, meta, mt);
        return true;
    }
    return false;
}

extern "C" {

void emu_counters(uint64_t* out, int reset) { for (int i = 0; i < 8; ++i) { out[i] = g_cnt[i]; if (reset) g_cnt[i] = 0; } }



// Maximal exact matches of one read: out gets 4 x u32 per match (start, end, lo, cnt), sorted by
// end.  Returns the number of matches.
int emu_sweep(const EmuIndex* ei, const uint32_t* words, uint32_t L, uint32_t* out, uint32_t cap, uint64_t* n_steps, uint32_t seed_K,
              const void* seed_tab) {
    HostIndex ix{(const Half*)ei->fwd, (const Half*)ei->rev, ei->sa, ei->text, {}, ei->n_bases};
    for (int c = 0; c < 4; ++c) { ix.meta.C[c] = ei->C[c]; ix.meta.cnt[c] = ei->cnt[c]; }
    ix.meta.prim_f = ei->prim_f; ix.meta.prim_r = ei->prim_r; ix.meta.n_rows = ei->n_rows;
    SweepCtx ctx; ctx.words = words; ctx.L = L; ctx.K = seed_tab ? seed_K : 0; ctx.use_uniq = ei->uniq_mode != 0; ctx.use_uniq_back = ei->uniq_mode == 2 && ei->isa != nullptr;
    Sweeper<SweepCtx> sw;
    uint64_t steps = 0;
    for (;;) {
        if (!sw.next(ctx, ix.meta)) break;
        if (sw.pending_seed()) {
            const SeedEntry* t = (const SeedEntry*)seed_tab;
            g_cnt[6]++;                                                      // [6] seed-table fetches
            sw.consume_seed(ctx, ix.meta, t[sw.P0]);
            ++steps;
            continue;
        }
        if (serve_uniq(sw, ctx, ix.meta, ei, words)) { ++steps; continue; }
        const bool rev = sw.on_reverse();
        const Half* bk = rev ? ix.rev : ix.fwd;
        auto load = [&](uint64_t idx) { return bk[idx]; };
        StepOut r = step_single(load, sw.P0, sw.P0 + sw.cnt, sw.ch, ix.meta.C[sw.ch], rev ? ix.meta.prim_r : ix.meta.prim_f);
        g_cnt[4 + (sw.mode == M_FWD ? 0 : 1)]++;                             // [4] FWD [5] WALK steps
        { uint32_t b0, r0, b1, r1; split192(sw.P0, b0, r0); split192(sw.P0 + sw.cnt, b1, r1); g_cnt[7] += (b0 != b1); }   // [7] steps touching two buckets
        sw.consume(ctx, ix.meta, r);
        ++steps;
    }
    sort_mems(ctx.mems);
    if (n_steps) *n_steps = steps;
    uint32_t n = (uint32_t)ctx.mems.size();
    for (uint32_t k = 0; k < n && k < cap; ++k) {
        out[4 * k] = ctx.mems[k].se & 0xFFFF; out[4 * k + 1] = ctx.mems[k].se >> 16;
        out[4 * k + 2] = ctx.mems[k].lo; out[4 * k + 3] = ctx.mems[k].cnt;
    }
    return (int)n;
}

// Full per-read pipeline (sweep + selection).  out: 6 x u32 per record.  Returns #records, or
// -1 if the reference would raise, -2 if the read is shorter than K.
int emu_smem(const EmuIndex* ei, int method, const uint32_t* words, uint32_t L, uint32_t min_len, uint32_t K,
             const uint32_t* lut, uint32_t n_levels, const uint32_t* level_sizes, const double* coef,
             const double* intercept, uint32_t* out, uint32_t cap, uint32_t seed_K, const void* seed_tab, uint32_t n_none,
             const uint32_t* none_rows) {
    HostIndex ix{(const Half*)ei->fwd, (const Half*)ei->rev, ei->sa, ei->text, {}, ei->n_bases};
    for (int c = 0; c < 4; ++c) { ix.meta.C[c] = ei->C[c]; ix.meta.cnt[c] = ei->cnt[c]; }
    ix.meta.prim_f = ei->prim_f; ix.meta.prim_r = ei->prim_r; ix.meta.n_rows = ei->n_rows;
    if (method != 0 && L < K) return -2;
    SweepCtx ctx; ctx.words = words; ctx.L = L; ctx.K = seed_tab ? seed_K : 0; ctx.use_uniq = ei->uniq_mode != 0; ctx.use_uniq_back = ei->uniq_mode == 2 && ei->isa != nullptr;
    Sweeper<SweepCtx> sw;
    for (;;) {
        if (!sw.next(ctx, ix.meta)) break;
        if (sw.pending_seed()) {
            sw.consume_seed(ctx, ix.meta, ((const SeedEntry*)seed_tab)[sw.P0]);
            continue;
        }
        if (serve_uniq(sw, ctx, ix.meta, ei, words)) continue;
        const bool rev = sw.on_reverse();
        const Half* bk = rev ? ix.rev : ix.fwd;
        auto load = [&](uint64_t idx) { return bk[idx]; };
        StepOut r = step_single(load, sw.P0, sw.P0 + sw.cnt, sw.ch, ix.meta.C[sw.ch], rev ? ix.meta.prim_r : ix.meta.prim_f);
        sw.consume(ctx, ix.meta, r);
    }
    sort_mems(ctx.mems);
    std::vector<uint32_t> recs;
    SelCtx sc;
    sc.ix = &ix; sc.words = words; sc.L = L; sc.K = K; sc.n_mems = (uint32_t)ctx.mems.size(); sc.min_len = min_len;
    sc.mems = ctx.mems.data(); sc.method = method; sc.lut = lut; sc.out = &recs;
    sc.seed_K = seed_tab ? seed_K : 0; sc.seed_tab = (const U4*)seed_tab;
    memset(&sc.rmi, 0, sizeof(sc.rmi));
    if (method == 2) {
        sc.rmi.K = K; sc.rmi.n_levels = n_levels; sc.rmi.coef = coef; sc.rmi.intercept = intercept; sc.rmi.stride = 1;
        uint32_t off = 0;
        for (uint32_t l = 0; l < n_levels; ++l) { sc.rmi.level_size[l] = level_sizes[l]; sc.rmi.level_off[l] = off; off += level_sizes[l]; }
        if (none_rows && n_none) rmi_set_none_rows(sc.rmi, none_rows, n_none, ei->n_rows);
    }
    if (method == 0) Selector<SelCtx>::run_bwa(sc);
    else Selector<SelCtx>::run_seeded(sc);
    if (sc.raised) return -1;
    uint32_t n = (uint32_t)(recs.size() / 6);
    for (uint32_t k = 0; k < n * 6 && k < cap * 6; ++k) out[k] = recs[k];
    return (int)n;
}

// Dense LUT exactly as the device builder computes it: backward search per code.
void emu_lut_build(const EmuIndex* ei, uint32_t K, uint32_t* table) {
    const Half* fwd = (const Half*)ei->fwd;
    auto load = [&](uint64_t idx) { return fwd[idx]; };
    uint64_t ncodes = 1ull << (2 * K);
    for (uint64_t code = 0; code < ncodes; ++code) {
        uint32_t lo = 0, cnt = ei->n_rows;
        for (uint32_t t = 0; t < K && cnt; ++t) {
            uint32_t c = (uint32_t)(code >> (2 * t)) & 3u;   // last base first
            StepOut r = step_single(load, lo, lo + cnt, c, ei->C[c], ei->prim_f);
            lo = r.lo_new; cnt = r.cnt_new;
        }
        table[2 * code] = lo; table[2 * code + 1] = cnt;
    }
}

// Seed table exactly as the device builder computes it (k_seed_build): per code the k-mer's rows on the text index
// and the reversed k-mer's rows on the reversed-text index.
void emu_seed_build(const EmuIndex* ei, uint32_t K, uint32_t* table) {
    const Half* fwd = (const Half*)ei->fwd;
    const Half* rev = (const Half*)ei->rev;
    auto lf = [&](uint64_t idx) { return fwd[idx]; };
    auto lr = [&](uint64_t idx) { return rev[idx]; };
    uint64_t ncodes = 1ull << (2 * K);
    for (uint64_t code = 0; code < ncodes; ++code) {
        uint32_t lo = 0, cnt = ei->n_rows, rlo = 0, rcnt = ei->n_rows;
        for (uint32_t t = 0; t < K; ++t) {                                 // through empty intervals too: lo stays the insertion point
            uint32_t c = (uint32_t)(code >> (2 * t)) & 3u;                 // last base first: backward search on the text
            StepOut r = step_single(lf, lo, lo + cnt, c, ei->C[c], ei->prim_f);
            lo = r.lo_new; cnt = r.cnt_new;
        }
        for (uint32_t t = 0; t < K && rcnt && cnt; ++t) {
            uint32_t c = (uint32_t)(code >> (2 * (K - 1 - t))) & 3u;       // first base first: backward search of the reversed k-mer
            StepOut r = step_single(lr, rlo, rlo + rcnt, c, ei->C[c], ei->prim_r);
            rlo = r.lo_new; rcnt = r.cnt_new;
        }
        table[4 * code] = lo; table[4 * code + 1] = cnt; table[4 * code + 2] = cnt ? rlo : 0; table[4 * code + 3] = 0;
    }
}

// rows -> 1-based positions through lf_single and a sampled suffix array (what k_locate_sampled runs)
void emu_locate(const EmuIndex* ei, uint32_t S, uint64_t n, const uint32_t* rows, uint32_t* pos) {
    const Half* fwd = (const Half*)ei->fwd;
    auto load = [&](uint64_t idx) { return fwd[idx]; };
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t r = rows[i], t = 0;
        for (;;) {
            if (r == ei->prim_f) { pos[i] = 1u + t; break; }
            if (r % S == 0u) { pos[i] = ei->sa[r] + t; break; }
            r = lf_single(load, r, ei->C, ei->prim_f);
            ++t;
        }
    }
}

// get_suffix_rmi for one code
int emu_rmi_lookup(const EmuIndex* ei, uint32_t K, uint32_t n_levels, const uint32_t* level_sizes, const double* coef,
                   const double* intercept, uint64_t code, double* pred, int64_t* lo, int64_t* hi) {
    RmiModel m; memset(&m, 0, sizeof(m));
    m.K = K; m.n_levels = n_levels; m.coef = coef; m.intercept = intercept; m.stride = 1;
    uint32_t off = 0;
    for (uint32_t l = 0; l < n_levels; ++l) { m.level_size[l] = level_sizes[l]; m.level_off[l] = off; off += level_sizes[l]; }
    auto sa = [&](uint64_t r) { return ei->sa[r]; };
    auto tx = [&](uint64_t w) { return ei->text[w]; };
    SaTextProbe<decltype(sa), decltype(tx)> pr{sa, tx};
    RmiTable<decltype(pr)> t{pr, (int64_t)ei->n_rows, (int64_t)ei->n_bases, K, false};
    t.lookup(m, code, *pred, *lo, *hi);
    return t.raised ? -1 : 0;
}

// the same lookup through the resumable single-probe-site machine (what the kernels run)
int emu_rmi_search(const EmuIndex* ei, uint32_t K, uint32_t n_levels, const uint32_t* level_sizes, const double* coef,
                   const double* intercept, uint64_t code, int64_t* lo, int64_t* hi, uint32_t* n_probes) {
    RmiModel m; memset(&m, 0, sizeof(m));
    m.K = K; m.n_levels = n_levels; m.coef = coef; m.intercept = intercept; m.stride = 1;
    uint32_t off = 0;
    for (uint32_t l = 0; l < n_levels; ++l) { m.level_size[l] = level_sizes[l]; m.level_off[l] = off; off += level_sizes[l]; }
    auto sa = [&](uint64_t r) { return ei->sa[r]; };
    auto tx = [&](uint64_t w) { return ei->text[w]; };
    SaTextProbe<decltype(sa), decltype(tx)> pr{sa, tx};
    RmiSearch rs;
    rs.begin(m, code, (int64_t)ei->n_rows, (int64_t)ei->n_bases);
    uint32_t np = 0;
    while (rs.pending()) {
        int64_t s; uint64_t c64;
        pr(rs.row(), s, c64);
        rs.feed(s, c64);
        ++np;
    }
    *lo = rs.out_lo; *hi = rs.out_hi; *n_probes = np;
    return rs.raised ? -1 : 0;
}

// the error-bounded fast search; *hazard = 1 when it defers to the literal search
int emu_rmi_fast(const EmuIndex* ei, uint32_t K, uint32_t n_levels, const uint32_t* level_sizes, const double* coef,
                 const double* intercept, uint32_t n_none, const uint32_t* none_rows, uint64_t code, int64_t* lo, int64_t* hi,
                 uint32_t* hazard, uint32_t* n_probes) {
    RmiModel m; memset(&m, 0, sizeof(m));
    m.K = K; m.n_levels = n_levels; m.coef = coef; m.intercept = intercept; m.stride = 1;
    uint32_t off = 0;
    for (uint32_t l = 0; l < n_levels; ++l) { m.level_size[l] = level_sizes[l]; m.level_off[l] = off; off += level_sizes[l]; }
    rmi_set_none_rows(m, none_rows, n_none, ei->n_rows);
    auto sa = [&](uint64_t r) { return ei->sa[r]; };
    auto tx = [&](uint64_t w) { return ei->text[w]; };
    SaTextProbe<decltype(sa), decltype(tx)> pr{sa, tx};
    *lo = 0; *hi = -1;
    const bool ok = rmi_fast_lookup(pr, m, code, ei->n_rows, (int64_t)ei->n_bases, *lo, *hi, n_probes);
    *hazard = ok ? 0u : 1u;
    return 0;
}

// the probe-free error-bounded lookup (rmi_arith_lookup): true bounds by a plain backward search continued through empty
// intervals, then the arithmetic replay of the exponential phase; *hazard = 1 when it defers to the literal search
int emu_rmi_arith(const EmuIndex* ei, uint32_t K, uint32_t n_levels, const uint32_t* level_sizes, const double* coef,
                  const double* intercept, uint32_t n_none, const uint32_t* none_rows, uint64_t code, int64_t* lo, int64_t* hi,
                  uint32_t* hazard) {
    RmiModel m; memset(&m, 0, sizeof(m));
    m.K = K; m.n_levels = n_levels; m.coef = coef; m.intercept = intercept; m.stride = 1;
    uint32_t off = 0;
    for (uint32_t l = 0; l < n_levels; ++l) { m.level_size[l] = level_sizes[l]; m.level_off[l] = off; off += level_sizes[l]; }
    rmi_set_none_rows(m, none_rows, n_none, ei->n_rows);
    const Half* fwd = (const Half*)ei->fwd;
    auto load = [&](uint64_t idx) { return fwd[idx]; };
    uint32_t A = 0, cnt = ei->n_rows;
    for (uint32_t t = 0; t < K; ++t) {
        const uint32_t c = (uint32_t)(code >> (2 * t)) & 3u;
        const StepOut r = step_single(load, A, A + cnt, c, ei->C[c], ei->prim_f);
        A = r.lo_new; cnt = r.cnt_new;
    }
    *lo = 0; *hi = -1;
    const bool ok = rmi_arith_lookup(m, RmiGallop::predicted_row(m, code, ei->n_rows), A, cnt, ei->n_rows, *lo, *hi);
    *hazard = ok ? 0u : 1u;
    return 0;
}

// The BWA-SMEM picks as the sweep's hand-over makes them (maximum of bwa_pick_key over the ordered list, continue from the
// pick's end) on a list of n <= 32 (start, end) pairs: the pick mask, bit = position in the list.
uint32_t emu_bwa_picks(uint32_t n, const uint32_t* starts, const uint32_t* ends) {
    uint32_t picks = 0;
    if (n == 0 || n > 32) return 0;
    for (uint32_t p = 0; p < ends[n - 1];) {
        uint32_t best = 0;
        for (uint32_t k = 0; k < n; ++k) { const uint32_t key = bwa_pick_key(starts[k], ends[k], k, p); if (key > best) best = key; }
        const uint32_t b = bwa_pick_of(best);
        picks |= 1u << b;
        p = ends[b];
    }
    return picks;
}

// closed-form rmi_arith_lookup against the probe-by-probe loops on n random (start, A, cnt) triples of a table of n_rows
// rows with the given None rows; returns the number of disagreements (outcome, bounds)
uint64_t emu_rmi_arith_fuzz(uint32_t n_rows, uint32_t n_none, const uint32_t* none_rows, uint64_t n, uint64_t seed) {
    RmiModel m; memset(&m, 0, sizeof(m));
    rmi_set_none_rows(m, none_rows, n_none, n_rows);
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1, bad = 0;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (uint64_t t = 0; t < n; ++t) {
        const uint32_t mode = (uint32_t)(rnd() % 6);
        uint32_t A = (uint32_t)(rnd() % ((uint64_t)n_rows + 1));
        uint32_t cnt = (uint32_t)(rnd() % 5 == 0 ? 0 : rnd() % ((uint64_t)n_rows - A + 1));
        if (mode == 1) cnt = (uint32_t)(rnd() % 4) < n_rows - A ? (uint32_t)(rnd() % 4) : 0;
        int64_t start = (int64_t)(rnd() % n_rows);
        if (mode == 2) start = (int64_t)A + (int64_t)(rnd() % 9) - 4;                 // near the lower bound
        if (mode == 3) start = (int64_t)A + cnt + (int64_t)(rnd() % 9) - 4;           // near the upper bound
        if (mode == 4) start = (int64_t)(rnd() % 3);                                  // table start
        if (mode == 5) start = (int64_t)n_rows - 1 - (int64_t)(rnd() % 3);            // table end
        if (start < 0) start = 0;
        if (start >= (int64_t)n_rows) start = -1;                                     // also: prediction outside the table
        int64_t l1 = 7, h1 = 7, l2 = 7, h2 = 7;
        const bool a = rmi_arith_lookup(m, start, A, cnt, n_rows, l1, h1);
        const bool b = rmi_arith_lookup_loops(m, start, A, cnt, n_rows, l2, h2);
        if (a != b || l1 != l2 || h1 != h2) ++bad;
    }
    return bad;
}

// The hazard codes of a model (rmi_code_is_hazard over every K-mer code, true bounds by a plain backward search continued
// through empty intervals, as emu_rmi_arith): writes up to cap codes, returns how many there are.  What k_rmi_hazard_scan computes.
uint64_t emu_rmi_hazards(const EmuIndex* ei, uint32_t K, uint32_t n_levels, const uint32_t* level_sizes, const double* coef,
                         const double* intercept, uint32_t n_none, const uint32_t* none_rows, uint32_t* out, uint64_t cap) {
    RmiModel m; memset(&m, 0, sizeof(m));
    m.K = K; m.n_levels = n_levels; m.coef = coef; m.intercept = intercept; m.stride = 1;
    uint32_t off = 0;
    for (uint32_t l = 0; l < n_levels; ++l) { m.level_size[l] = level_sizes[l]; m.level_off[l] = off; off += level_sizes[l]; }
    rmi_set_none_rows(m, none_rows, n_none, ei->n_rows);
    const Half* fwd = (const Half*)ei->fwd;
    auto load = [&](uint64_t idx) { return fwd[idx]; };
    uint64_t n = 0;
    for (uint64_t code = 0; code < (1ull << (2 * K)); ++code) {
        uint32_t A = 0, cnt = ei->n_rows;
        for (uint32_t t = 0; t < K; ++t) {
            const uint32_t c = (uint32_t)(code >> (2 * t)) & 3u;
            const StepOut r = step_single(load, A, A + cnt, c, ei->C[c], ei->prim_f);
            A = r.lo_new; cnt = r.cnt_new;
        }
        if (rmi_code_is_hazard(m, code, A, cnt, ei->n_rows)) { if (n < cap) out[n] = (uint32_t)code; ++n; }
    }
    return n;
}

int emu_hz_build(const uint32_t* codes, uint64_t n, uint32_t* slots, uint32_t n_slots) { return hz_build(codes, n, slots, n_slots) ? 1 : 0; }

int emu_hz_contains(const uint32_t* slots, uint32_t n_slots, uint32_t code) {
    return hz_contains([slots](uint32_t h) { return slots[h]; }, n_slots - 1u, code) ? 1 : 0;
}

// rmi_read_hazard_free on a packed read, as k_rmi_prefilter calls it
int emu_read_hazard_free(const uint32_t* words, uint32_t L, uint32_t K, const uint32_t* slots, uint32_t n_slots) {
    return rmi_read_hazard_free([words](uint32_t w) { return words[w]; }, L, K, [slots](uint32_t h) { return slots[h]; }, n_slots - 1u) ? 1 : 0;
}

// kmer_code of the window at base p (what the selection kernels look up), for the cross-check of the rolled codes
uint64_t emu_kmer_code(const uint32_t* words, uint32_t p, uint32_t K) {
    return kmer_code([words](uint64_t w) { return words[w]; }, p, K);
}

}  // extern "C"
