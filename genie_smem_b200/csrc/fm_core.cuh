// FM-index rank arithmetic shared by every kernel (and by the host-compiled logic test).
//
// One rank bucket = 64 bytes = 16 words:
//   w[0..3]           raw counts of A,C,G,T in rows [0, 192*b)   ('$' slot counted as A)
//   w[4k], w[4k+1]    low-bit  plane of symbols 64(k-1) .. 64(k-1)+63     (k = 1..3)
//   w[4k+2], w[4k+3]  high-bit plane of the same symbols
// so a rank query is ONE aligned 64-byte fetch: four lanes ("a quad") load 16 bytes each with
// a single 128-bit load; lane 0 holds the checkpoint, lanes 1..3 popcount 64 symbols each.
// This replaces the reference's n x 5 inclusive occurrence matrix (ExactMatch.py:70-90) and the
// two list reads per step of exact_match_back_prop (ExactMatch.py:140-145).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GSM_HD __host__ __device__ __forceinline__
#else
#define GSM_HD inline
#endif

namespace gsm {

struct U4 {
    uint32_t x, y, z, w;
};

GSM_HD uint32_t popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

// p / 192 and p % 192 for 32-bit row numbers
GSM_HD void split192(uint32_t p, uint32_t& bucket, uint32_t& r) {
#if defined(__CUDA_ARCH__)
    bucket = __umulhi(p >> 6, 0xAAAAAAABu) >> 1;
#else
    bucket = (p >> 6) / 3u;
#endif
    r = p - bucket * 192u;
}

// Per-lane contribution of one 16-byte part of a bucket to a rank query at in-bucket offset r:
// returns (#symbols == c) | (#symbols < c) << 8 among the first r symbols, counting only the 64
// symbols this part (ql = 1..3) holds.  ql == 0 (the checkpoint part) contributes 0.
GSM_HD uint32_t part_counts(const U4& v, uint32_t r, uint32_t c, uint32_t ql) {
    int nsym = (int)r - 64 * ((int)ql - 1);
    if (ql == 0 || nsym <= 0) return 0u;
    uint32_t m0, m1;
    if (nsym >= 64) { m0 = 0xFFFFFFFFu; m1 = 0xFFFFFFFFu; }
    else if (nsym >= 32) { m0 = 0xFFFFFFFFu; m1 = (nsym == 32) ? 0u : ((1u << (nsym - 32)) - 1u); }
    else { m0 = (1u << nsym) - 1u; m1 = 0u; }
    const uint32_t fl = (c & 1u) ? 0u : 0xFFFFFFFFu;   // flip so that "bit == c's bit" reads as 1
    const uint32_t fh = (c & 2u) ? 0u : 0xFFFFFFFFu;
    const uint32_t X = (c >= 2u) ? 0xFFFFFFFFu : 0u;
    const uint32_t Y = (c == 3u) ? 0xFFFFFFFFu : 0u;
    const uint32_t Z = (c != 0u) ? 0xFFFFFFFFu : 0u;
    const uint32_t L0 = v.x, L1 = v.y, H0 = v.z, H1 = v.w;
    uint32_t eq = popc32((L0 ^ fl) & (H0 ^ fh) & m0) + popc32((L1 ^ fl) & (H1 ^ fh) & m1);
    // symbols < c:  c=1: ~H&~L   c=2: ~H   c=3: ~H|~L   c=0: none
    uint32_t lt = popc32(((~H0 & (~L0 | X)) | (~L0 & Y)) & m0 & Z) + popc32(((~H1 & (~L1 | X)) | (~L1 & Y)) & m1 & Z);
    return eq | (lt << 8);
}

// Checkpoint part: counts of c and of symbols < c before the bucket.
GSM_HD void header_counts(const U4& v, uint32_t c, uint32_t& h_eq, uint32_t& h_lt) {
    h_eq = c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w;
    h_lt = c == 0 ? 0u : c == 1 ? v.x : c == 2 ? v.x + v.y : v.x + v.y + v.z;
}

// Result of one extension step on an index with rows [P0, P1) and symbol c:
//   lo_new  = C[c] + rank(c, P0)                 new interval start on THIS index
//   cnt_new = rank(c, P1) - rank(c, P0)          new interval size (0 => no match)
//   lt_add  = ['$' in rows [P0,P1)] + sum_{b<c} (rank(b,P1) - rank(b,P0))
//             = how far the interval start moves on the OTHER (reverse-text) index
struct StepOut {
    uint32_t lo_new, cnt_new, lt_add;
};

// Combine raw totals (checkpoint + in-bucket, '$' still counted as A) into a step result.
GSM_HD StepOut finish_step(uint32_t eq0, uint32_t lt0, uint32_t eq1, uint32_t lt1, uint32_t P0, uint32_t P1,
                           uint32_t c, uint32_t Cc, uint32_t primary) {
    const uint32_t a0 = P0 > primary ? 1u : 0u;   // the fake A of the '$' slot lies below P0
    const uint32_t a1 = P1 > primary ? 1u : 0u;
    const uint32_t inrange = a1 - a0;
    StepOut o;
    if (c == 0) {
        o.lo_new = Cc + eq0 - a0;
        o.cnt_new = (eq1 - eq0) - inrange;
        o.lt_add = inrange;
    } else {
        o.lo_new = Cc + eq0;
        o.cnt_new = eq1 - eq0;
        o.lt_add = lt1 - lt0;   // raw lt counts include the fake A exactly when '$' is in range
    }
    return o;
}

// Whole-bucket rank by ONE thread (used by the selection kernels' rare paths, the LUT builder
// and the host-compiled logic test).  bk = bucket array as 16-byte parts.
template <typename LoadU4>
GSM_HD StepOut step_single(LoadU4 load, uint32_t P0, uint32_t P1, uint32_t c, uint32_t Cc, uint32_t primary) {
    uint32_t tot[2][2];
    const uint32_t P[2] = {P0, P1};
    for (int e = 0; e < 2; ++e) {
        uint32_t b, r;
        split192(P[e], b, r);
        uint32_t h_eq, h_lt;
        U4 h = load((uint64_t)b * 4);
        header_counts(h, c, h_eq, h_lt);
        uint32_t acc = 0;
        for (uint32_t ql = 1; ql <= 3; ++ql) {
            if ((int)r - 64 * ((int)ql - 1) <= 0) break;
            U4 v = load((uint64_t)b * 4 + ql);
            acc += part_counts(v, r, c, ql);
        }
        tot[e][0] = h_eq + (acc & 0xFFu);
        tot[e][1] = h_lt + ((acc >> 8) & 0xFFu);
    }
    return finish_step(tot[0][0], tot[0][1], tot[1][0], tot[1][1], P0, P1, c, Cc, primary);
}

// Packed sequences (reads and text) are MSB-first: base i lives in bits [30-2(i%16), 32-2(i%16))
// of word i/16.
GSM_HD uint32_t base_msb(const uint32_t* words, uint32_t pos) { return (words[pos >> 4] >> (30u - 2u * (pos & 15u))) & 3u; }

// MSB-first k-mer code (LUT.convert_seq_to_num, reference SMEM/LUT.py:37-48) of K <= 32 bases
// starting at base p.  The array must be readable two words past the last base.
template <typename LoadW>
GSM_HD uint64_t kmer_code(LoadW load, uint64_t p, uint32_t K) {
    uint64_t w = p >> 4;
    uint32_t sh = 2u * (uint32_t)(p & 15u);
    uint64_t hi = ((uint64_t)load(w) << 32) | (uint64_t)load(w + 1);
    uint64_t v = hi << sh;
    if (sh) v |= (uint64_t)load(w + 2) >> (32u - sh);
    return v >> (64u - 2u * K);
}

}  // namespace gsm
