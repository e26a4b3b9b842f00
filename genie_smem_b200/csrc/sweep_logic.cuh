// Per-read control of the SMEM sweep kernel, written once for device and for the host-compiled
// logic test (tests/emu).  It enumerates ALL super-maximal exact matches of a read with a
// bidirectional FM index, the way the reference's forward_extension / backward_extension pair
// (SMEM/SMEM.py:425-443, 389-423) would if each extension cost one FM step instead of a full
// restart of exact_match_back_prop:
//
//   sweep(x):  forward phase  -- append q[x], q[x+1], ... on the reverse-text index while the
//                                pattern q[x:j] still occurs (SMEM.py:431-440); remember (j,
//                                interval) whenever the occurrence count is about to change:
//                                only those ends can be right-maximal;
//              backward phase -- prepend q[x-1], q[x-2], ... to all remembered candidates in
//                                lock step, longest first (SMEM.py:396-416); a candidate that
//                                dies while no longer one survives is a maximal match;
//              x = end of the longest forward match; repeat until x == L.
//
// With LS[j] = leftmost start of a match ending at j the emitted set is
// {(LS[j], j) : j = L or LS[j+1] > LS[j]}; it determines LS[] and F() completely, and the three
// reference entry points are integer selections over it (select_logic.cuh).
//
// One call of prepare()/consume() pair = at most one FM extension step = two bucket fetches.
#pragma once
#include "fm_core.cuh"

namespace gsm {

struct IndexMeta {
    uint32_t C[4];      // first row per leading base
    uint32_t cnt[4];    // occurrences per base
    uint32_t prim_f, prim_r;
    uint32_t n_rows;
};

enum SweepPhase : int { PH_FETCH = 0, PH_FWD = 1, PH_BWD = 2, PH_DONE = 3 };

// A maximal-match record as staged per read: 16 bytes.
//   x = start | end << 16, y = SA lo, z = SA count, w = sweep ordinal
struct MemEntry {
    uint32_t se, lo, cnt, sweep;
};

// Ctx must provide:
//   bool     fetch(uint32_t& rid, uint32_t& L)          next read (loads its bases), false when none
//   uint32_t base(uint32_t pos)                         2-bit base of the current read
//   void     cand_put(uint32_t i, uint32_t j, uint32_t lo, uint32_t cnt)
//   void     cand_get(uint32_t i, uint32_t& j, uint32_t& lo, uint32_t& cnt)
//   void     cand_sync()                                make candidate writes visible to the quad
//   void     emit(uint32_t idx, MemEntry e)             stage maximal match #idx of this read
//   void     finish(uint32_t rid, uint32_t n_mems)      read complete
template <typename Ctx>
struct Sweeper {
    int ph = PH_FETCH;
    uint32_t rid = 0, L = 0;
    uint32_t x = 0;          // sweep start
    uint32_t j = 0;          // forward: pattern is q[x:j]
    uint32_t k = 0, l = 0, s = 0;   // forward: rows on the text index / reverse index / count
    uint32_t ncand = 0;
    int32_t i = 0;           // backward: position being prepended
    uint32_t t = 0, base_i = 0, top = 0, w = 0;
    uint32_t lastkept = 0, last_start = 0;
    uint32_t F = 0;
    uint32_t n_mems = 0, sweep_id = 0;
    // pending step
    uint32_t cur_j = 0, cur_lo = 0, cur_cnt = 0;

    GSM_HD void push(Ctx& c, uint32_t jj, uint32_t lo, uint32_t cnt) { c.cand_put(ncand++, jj, lo, cnt); }

    GSM_HD void begin_bwd(Ctx& c) {
        F = j;
        base_i = 0; top = ncand; t = ncand; w = ncand;
        i = (int32_t)x - 1;
        last_start = 0xFFFFFFFFu;
        lastkept = 0;
        ph = PH_BWD;
        c.cand_sync();
    }

    GSM_HD void begin_fwd(Ctx& c, const IndexMeta& m) {
        for (;;) {
            uint32_t b = c.base(x);
            ncand = 0;
            k = m.C[b]; l = m.C[b]; s = m.cnt[b];
            j = x + 1;
            if (s == 0) {            // base absent from the text: no match covers x (outside the
                x++;                 // reference's domain); skip it
                if (x >= L) { c.finish(rid, n_mems); ph = PH_FETCH; return; }
                continue;
            }
            if (j == L) { push(c, j, k, s); begin_bwd(c); return; }
            ph = PH_FWD;
            return;
        }
    }

    GSM_HD void fail_cand(Ctx& c) {
        // candidate (cur_j, cur_lo, cur_cnt) cannot be extended to position i: match q[i+1 : cur_j]
        if (w == top && (uint32_t)(i + 1) < last_start) {
            MemEntry e;
            e.se = (uint32_t)(i + 1) | (cur_j << 16);
            e.lo = cur_lo; e.cnt = cur_cnt; e.sweep = sweep_id;
            c.emit(n_mems++, e);
            last_start = (uint32_t)(i + 1);
        }
    }

    // Runs the zero-cost transitions.  Returns true when an FM step is pending; then (P0, P1, ch,
    // on_reverse) describe it.  Returns false only when this quad has no more reads.
    GSM_HD bool prepare(Ctx& c, const IndexMeta& m, uint32_t& P0, uint32_t& P1, uint32_t& ch, bool& on_reverse) {
        for (;;) {
            if (ph == PH_DONE) return false;
            if (ph == PH_FETCH) {
                if (!c.fetch(rid, L)) { ph = PH_DONE; return false; }
                n_mems = 0; sweep_id = 0; x = 0;
                if (L == 0) { c.finish(rid, 0); continue; }
                begin_fwd(c, m);
                continue;
            }
            if (ph == PH_FWD) {
                P0 = l; P1 = l + s; ch = c.base(j); on_reverse = true;
                return true;
            }
            // PH_BWD
            if (t > base_i) {
                c.cand_get(t - 1, cur_j, cur_lo, cur_cnt);
                if (i >= 0) {
                    P0 = cur_lo; P1 = cur_lo + cur_cnt; ch = c.base((uint32_t)i); on_reverse = false;
                    return true;
                }
                fail_cand(c);     // ran off the left end of the read
                t--;
                continue;
            }
            // round finished
            if (w == top) {       // nobody survived: the sweep is over
                x = F;
                sweep_id++;
                if (x >= L) { c.finish(rid, n_mems); ph = PH_FETCH; continue; }
                begin_fwd(c, m);
                continue;
            }
            base_i = w; t = top; w = top; i--; lastkept = 0;
            c.cand_sync();
        }
    }

    GSM_HD void consume(Ctx& c, const IndexMeta& m, const StepOut& r) {
        (void)m;
        if (ph == PH_FWD) {
            if (r.cnt_new != s) push(c, j, k, s);
            if (r.cnt_new == 0) { begin_bwd(c); return; }
            k += r.lt_add; l = r.lo_new; s = r.cnt_new; j++;
            if (j == L) { push(c, j, k, s); begin_bwd(c); }
            return;
        }
        // PH_BWD: candidate t-1 was extended with q[i]
        if (r.cnt_new == 0) {
            fail_cand(c);
        } else if (w == top || r.cnt_new != lastkept) {
            w--;
            c.cand_put(w, cur_j, r.lo_new, r.cnt_new);
            lastkept = r.cnt_new;
        }
        t--;
    }
};

}  // namespace gsm
