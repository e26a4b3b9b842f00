// Per-read control of the SMEM sweep kernel, written once for device and for the host-compiled
// logic test (tests/emu).  It enumerates ALL maximal exact matches of a read with a bidirectional
// FM index, the way the reference's forward_extension / backward_extension pair
// (SMEM/SMEM.py:425-443, 389-423) would if each extension cost one FM step instead of a full
// restart of exact_match_back_prop:
//
//   sweep(x):  FWD   append q[x], q[x+1], ... on the reverse-text index while q[x:j] still occurs
//                    (SMEM.py:431-440); remember (j, interval) whenever the occurrence count is
//                    about to change: only those ends can be right-maximal.
//              WALK  prepend q[x-1], q[x-2], ... to the LONGEST candidate until it dies
//                    (SMEM.py:396-416): its start s is LS[F(x)].  Every other candidate j has
//                    lb <= LS[j] <= s, where lb = (previous sweep's start) + 1 because that
//                    sweep's forward extension ended exactly at x.  If s == lb they all share
//                    LS = s and none of them is maximal: the sweep is over after the minimum
//                    possible number of steps (the common case inside an exact stretch).
//              LOCK  otherwise prepend to the remaining candidates in lock step, longest first;
//                    a candidate that dies while no longer one survives, at a start left of
//                    every match found so far, is maximal; a survivor whose count equals the
//                    previous (longer) survivor's is dropped (same occurrences => same LS).
//                    A single survivor continues in WALK mode (registers only).
//              x = F(x) = end of the longest forward match; repeat until x == L.
//
// With LS[j] = leftmost start of a match ending at j the emitted set is
// {(LS[j], j) : j = L or LS[j+1] > LS[j]}; it determines LS[] and F() completely, and the three
// reference entry points are integer selections over it (select_logic.cuh).
//
// One next()/consume() pair = one FM extension step = at most two bucket fetches.  The pending
// step's operands live in registers (P0, cnt, ch), so the top of the kernel loop is uniform.
#pragma once
#include "fm_core.cuh"

namespace gsm {

struct IndexMeta {
    uint32_t C[4];      // first row per leading base
    uint32_t cnt[4];    // occurrences per base
    uint32_t prim_f, prim_r;
    uint32_t n_rows;
};

// A maximal-match record as staged per read: 16 bytes.
//   se = start | end << 16, lo = SA lo, cnt = SA count, sweep = sweep ordinal
struct MemEntry {
    uint32_t se, lo, cnt, sweep;
};

enum SweepMode : int { M_FETCH = 0, M_FWD = 1, M_WALK = 2, M_LOCK = 3, M_DONE = 4 };

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
    int mode = M_FETCH;
    uint32_t rid = 0, L = 0;
    uint32_t x = 0, F = 0, lb = 0;
    uint32_t n_mems = 0, sweep_id = 0;
    // pending step: rows [P0, P0 + cnt) extended by base ch (FWD: on the reverse-text index)
    uint32_t P0 = 0, cnt = 0, ch = 0;
    uint32_t k = 0;            // FWD: rows of q[x:pos) on the text index start here
    uint32_t pos = 0;          // FWD: next base to append; WALK/LOCK: base being prepended
    uint32_t cur_j = 0;        // WALK/LOCK: end of the candidate being extended
    uint32_t ncand = 0;
    uint32_t first_walk = 0;
    uint32_t t = 0, base_i = 0, top = 0, w = 0, lastkept = 0, last_start = 0;

    GSM_HD bool pending() const { return mode == M_FWD || mode == M_WALK || mode == M_LOCK; }
    GSM_HD bool on_reverse() const { return mode == M_FWD; }

    GSM_HD void emit_match(Ctx& c, uint32_t start, uint32_t end, uint32_t lo, uint32_t n) {
        MemEntry e;
        e.se = start | (end << 16);
        e.lo = lo; e.cnt = n; e.sweep = sweep_id;
        c.emit(n_mems++, e);
        last_start = start;
    }

    // A read is in progress and x < L: start sweeps until one has a pending FM step or the read ends.
    GSM_HD void start_sweeps(Ctx& c, const IndexMeta& m) {
        for (;;) {
            const uint32_t b = c.base(x);
            ncand = 0;
            k = m.C[b]; P0 = m.C[b]; cnt = m.cnt[b];
            pos = x + 1;
            if (cnt == 0) {              // base absent from the text (outside the reference's domain): skip it
                lb = x + 1; x++; sweep_id++;
                if (x >= L) { c.finish(rid, n_mems); return; }
                continue;
            }
            if (pos < L) { ch = c.base(pos); mode = M_FWD; return; }
            c.cand_put(ncand++, pos, k, cnt);
            if (start_bwd(c)) return;
            if (x >= L) return;          // end_sweep() finished the read
        }
    }

    // mode M_FETCH doubles as "needs a transition": x >= L means no read is in progress.
    GSM_HD void end_sweep(Ctx& c) {
        lb = x + 1; x = F; sweep_id++;
        mode = M_FETCH;
        if (x >= L) c.finish(rid, n_mems);
    }

    // Forward phase over: pop the longest candidate and walk it left.  Returns true if a step is pending.
    GSM_HD bool start_bwd(Ctx& c) {
        c.cand_sync();
        ncand--;
        c.cand_get(ncand, cur_j, P0, cnt);
        F = cur_j;
        first_walk = 1;
        last_start = 0xFFFFFFFFu;
        if (x == 0) {                       // nothing to prepend: every candidate starts at 0, the longest wins
            emit_match(c, 0, cur_j, P0, cnt);
            end_sweep(c);
            return false;
        }
        pos = x - 1;
        ch = c.base(pos);
        mode = M_WALK;
        return true;
    }

    GSM_HD void start_lock(Ctx& c) {
        first_walk = 0;
        base_i = 0; top = ncand; t = ncand; w = ncand; lastkept = 0;
        pos = x - 1;
        ch = c.base(pos);
        c.cand_get(t - 1, cur_j, P0, cnt);
        mode = M_LOCK;
    }

    // Bring this quad to its next pending step.  Returns false only when there are no more reads.
    GSM_HD bool next(Ctx& c, const IndexMeta& m) {
        while (mode == M_FETCH) {
            if (x >= L) {                   // no read in progress (initial state: x == L == 0)
                if (!c.fetch(rid, L)) { mode = M_DONE; return false; }
                n_mems = 0; sweep_id = 0; x = 0; lb = 0;
                if (L == 0) { c.finish(rid, 0); continue; }
            }
            start_sweeps(c, m);
        }
        return mode != M_DONE;
    }

    // The walked candidate (cur_j, P0, cnt) cannot be extended beyond `start`.
    GSM_HD void walk_end(Ctx& c, uint32_t start) {
        if (start < last_start) emit_match(c, start, cur_j, P0, cnt);
        if (first_walk && start != lb && ncand != 0) start_lock(c);
        else end_sweep(c);
    }

    GSM_HD void consume(Ctx& c, const IndexMeta& m, const StepOut& r) {
        (void)m;
        if (mode != M_LOCK) {
            // FWD (append q[pos] to q[x:pos)) and WALK (prepend q[pos] to q[pos+1:cur_j)) share one hot path:
            // take the new interval, move one base, fetch it.  Only the ends of an extension branch.
            const bool fwd = mode == M_FWD;
            if (fwd && r.cnt_new != cnt) c.cand_put(ncand++, pos, k, cnt);     // count about to change: q[x:pos) is a candidate
            if (r.cnt_new != 0) {
                k += r.lt_add; P0 = r.lo_new; cnt = r.cnt_new;
                const uint32_t last = fwd ? L - 1u : 0u;
                if (pos != last) { pos += fwd ? 1u : 0xFFFFFFFFu; ch = c.base(pos); return; }
                if (fwd) { pos++; c.cand_put(ncand++, pos, k, cnt); start_bwd(c); return; }    // ran off the right end
                walk_end(c, 0u);                                                               // ran off the left end
                return;
            }
            if (fwd) start_bwd(c);
            else walk_end(c, pos + 1u);
            return;
        }
        // M_LOCK: candidate t-1 = (cur_j, P0, cnt) was extended with q[pos]
        if (r.cnt_new == 0) {
            if (w == top && pos + 1 < last_start) emit_match(c, pos + 1, cur_j, P0, cnt);
        } else if (w == top || r.cnt_new != lastkept) {
            w--;
            c.cand_put(w, cur_j, r.lo_new, r.cnt_new);
            lastkept = r.cnt_new;
        }
        t--;
        if (t == base_i) {                       // round finished
            if (w == top) { end_sweep(c); return; }
            c.cand_sync();
            if (pos == 0) {                      // survivors ran off the left end: the longest one is the match
                c.cand_get(top - 1, cur_j, P0, cnt);
                if (0 < last_start) emit_match(c, 0, cur_j, P0, cnt);
                end_sweep(c);
                return;
            }
            base_i = w; t = top; w = top; lastkept = 0;
            pos--;
            ch = c.base(pos);
            if (top - base_i == 1) {             // one survivor: finish it in registers
                c.cand_get(top - 1, cur_j, P0, cnt);
                mode = M_WALK;
                return;
            }
        }
        c.cand_get(t - 1, cur_j, P0, cnt);
    }
};

}  // namespace gsm
