// FM-index rank arithmetic shared by every kernel (and by the host-compiled logic test).
//
// One rank bucket = 64 bytes = two 32-byte halves; half g (g = 0, 1) holds
//   w[0], w[1]        raw counts of symbols 2g and 2g+1 in rows [0, 192*b)   ('$' slot counted as A)
//   w[2..4]           low-bit  plane of symbols 96g .. 96g+95 of the bucket (bit t of word m = symbol 96g+32m+t)
//   w[5..7]           high-bit plane of the same symbols
// so a rank query is ONE aligned 64-byte fetch: two lanes ("a pair") load one half each with a
// single 256-bit load, popcount at most 96 symbols each, and add their two checkpoint words; both
// lanes do the same amount of useful work.  This replaces the reference's n x 5 inclusive occurrence
// matrix (ExactMatch.py:70-90) and the two list reads per step of exact_match_back_prop
// (ExactMatch.py:140-145).
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

// one 32-byte half of a bucket
struct Half {
    uint32_t c0, c1;        // checkpoint counts of symbols 2g, 2g+1
    uint32_t l0, l1, l2;    // low-bit plane
    uint32_t h0, h1, h2;    // high-bit plane
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

// n low bits set, n in 0..32
GSM_HD uint32_t mask_low(uint32_t n) {
#if defined(__CUDA_ARCH__)
    uint32_t m;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(m) : "r"(0u), "r"(n));
    return m;
#else
    return n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u);
#endif
}

GSM_HD uint32_t clamp32(int v) { return (uint32_t)(v < 0 ? 0 : (v > 32 ? 32 : v)); }

// Per-symbol constants of a rank query, computed once per step.
struct SymK {
    uint32_t fl, fh;    // plane flips so that "bit == c's bit" reads as 1
    uint32_t X, Y;      // lt selector:  c=1: ~H&~L   c=2: ~H   c=3: ~H|~L
    uint32_t nz;        // all ones unless c == 0 (nothing is smaller than A)
};

GSM_HD SymK sym_consts(uint32_t c) {
    SymK k;
    k.fl = 0u - ((c & 1u) ^ 1u);
    k.fh = 0u - (((c >> 1) & 1u) ^ 1u);
    k.X = 0u - (c >> 1);
    k.Y = 0u - (c & (c >> 1));
    k.nz = 0u - (uint32_t)(c != 0u);
    return k;
}

// In-bucket contribution of half g to a rank query at in-bucket offset r (0..191):
// (#symbols == c) | (#symbols < c) << 8 among the symbols of this half that lie below r.
GSM_HD uint32_t half_counts(const Half& v, uint32_t r, const SymK& k, uint32_t g) {
    const int n = (int)r - 96 * (int)g;             // symbols of this half below r (may be <0 or >96)
    const uint32_t m0 = mask_low(clamp32(n)), m1 = mask_low(clamp32(n - 32)), m2 = mask_low(clamp32(n - 64));
    const uint32_t eq = popc32((v.l0 ^ k.fl) & (v.h0 ^ k.fh) & m0) + popc32((v.l1 ^ k.fl) & (v.h1 ^ k.fh) & m1) +
                        popc32((v.l2 ^ k.fl) & (v.h2 ^ k.fh) & m2);
    const uint32_t lt = popc32(((~v.h0 & (~v.l0 | k.X)) | (~v.l0 & k.Y)) & m0 & k.nz) +
                        popc32(((~v.h1 & (~v.l1 | k.X)) | (~v.l1 & k.Y)) & m1 & k.nz) +
                        popc32(((~v.h2 & (~v.l2 | k.X)) | (~v.l2 & k.Y)) & m2 & k.nz);
    return eq | (lt << 8);
}

// Checkpoint contribution of half g: its share of count(c) and of count(symbols < c) before the bucket.
GSM_HD void half_header(const Half& v, uint32_t c, uint32_t g, uint32_t& h_eq, uint32_t& h_lt) {
    const uint32_t mine = (c >> 1) == g ? ((c & 1u) ? v.c1 : v.c0) : 0u;
    // half 0 holds A,C: A < c for c >= 1, C < c for c >= 2; half 1 holds G,T: G < c only for c == 3
    const uint32_t lt0 = (c >= 1u ? v.c0 : 0u) + (c >= 2u ? v.c1 : 0u);
    const uint32_t lt1 = c == 3u ? v.c0 : 0u;
    h_eq = mine;
    h_lt = g == 0u ? lt0 : lt1;
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
// eq0/eq1 = raw count(c) below P0/P1, ltd = raw count(< c) in [P0, P1).
GSM_HD StepOut finish_step(uint32_t eq0, uint32_t eq1, uint32_t ltd, uint32_t P0, uint32_t P1, uint32_t c, uint32_t Cc,
                           uint32_t primary) {
    const uint32_t a0 = P0 > primary ? 1u : 0u;   // the fake A of the '$' slot lies below P0
    const uint32_t a1 = P1 > primary ? 1u : 0u;
    const uint32_t inrange = a1 - a0;
    const uint32_t isA = c == 0u ? 1u : 0u;
    StepOut o;
    o.lo_new = Cc + eq0 - (isA & a0);
    o.cnt_new = (eq1 - eq0) - (isA & inrange);
    o.lt_add = isA ? inrange : ltd;   // raw lt counts include the fake A exactly when '$' is in range
    return o;
}

// "equals c" / "less than c" bit masks of a whole bucket's 6 x 32 symbols (half 0: words 0..2, half 1: words 3..5)
GSM_HD void bucket_masks(const Half& h0, const Half& h1, const SymK& k, uint32_t (&E)[6], uint32_t (&T)[6]) {
    const uint32_t L[6] = {h0.l0, h0.l1, h0.l2, h1.l0, h1.l1, h1.l2}, H[6] = {h0.h0, h0.h1, h0.h2, h1.h0, h1.h1, h1.h2};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int w = 0; w < 6; ++w) {
        E[w] = (L[w] ^ k.fl) & (H[w] ^ k.fh);
        T[w] = ((~H[w] & (~L[w] | k.X)) | (~L[w] & k.Y)) & k.nz;
    }
}

// counts below in-bucket offset r (0..191): (#symbols == c) | (#symbols < c) << 16
GSM_HD uint32_t bucket_counts(const uint32_t (&E)[6], const uint32_t (&T)[6], uint32_t r) {
    uint32_t eq = 0, lt = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int w = 0; w < 6; ++w) {
        const uint32_t m = mask_low(clamp32((int)r - 32 * w));
        eq += popc32(E[w] & m);
        lt += popc32(T[w] & m);
    }
    return eq | (lt << 16);
}

// Whole-bucket rank by ONE thread (selection kernels, table builders, host logic test).
// load(i) returns the i-th 32-byte half of the bucket array.  Rows P0 and P1 usually lie in the same bucket (narrow
// intervals): its two halves are loaded once and the symbol masks built once for both offsets.
template <typename LoadHalf>
GSM_HD StepOut step_single(LoadHalf load, uint32_t P0, uint32_t P1, uint32_t c, uint32_t Cc, uint32_t primary) {
    const SymK k = sym_consts(c);
    uint32_t b0, r0, b1, r1;
    split192(P0, b0, r0);
    split192(P1, b1, r1);
    uint32_t E[6], T[6];
    const Half a0 = load((uint64_t)b0 * 2), a1 = load((uint64_t)b0 * 2 + 1);
    bucket_masks(a0, a1, k, E, T);
    const uint32_t acc0 = bucket_counts(E, T, r0);
    uint32_t e0, l0, e1, l1;
    half_header(a0, c, 0u, e0, l0);
    half_header(a1, c, 1u, e1, l1);
    const uint32_t eq0 = e0 + e1 + (acc0 & 0xFFFFu), lt0 = l0 + l1 + (acc0 >> 16);
    uint32_t eq1, lt1;
    if (b1 == b0) {
        const uint32_t acc1 = bucket_counts(E, T, r1);
        eq1 = e0 + e1 + (acc1 & 0xFFFFu); lt1 = l0 + l1 + (acc1 >> 16);
    } else {
        const Half c0 = load((uint64_t)b1 * 2), c1 = load((uint64_t)b1 * 2 + 1);
        bucket_masks(c0, c1, k, E, T);
        const uint32_t acc1 = bucket_counts(E, T, r1);
        half_header(c0, c, 0u, e0, l0);
        half_header(c1, c, 1u, e1, l1);
        eq1 = e0 + e1 + (acc1 & 0xFFFFu); lt1 = l0 + l1 + (acc1 >> 16);
    }
    return finish_step(eq0, eq1, lt1 - lt0, P0, P1, c, Cc, primary);
}

// LF mapping by ONE thread: row r -> the row of the suffix that starts one base earlier (suffix_array[LF(r)] =
// suffix_array[r] - 1).  One bucket gives both the BWT symbol at r and its rank.  Must not be called for r == primary
// (that row's BWT symbol is '$': its suffix starts at text position 0).  Used by the sampled-SA locate kernel.
template <typename LoadHalf>
GSM_HD uint32_t lf_single(LoadHalf load, uint32_t r, const uint32_t* C, uint32_t primary) {
    uint32_t b, off;
    split192(r, b, off);
    const Half h0 = load((uint64_t)b * 2), h1 = load((uint64_t)b * 2 + 1);
    const uint32_t g = off >= 96u ? 1u : 0u, o = off - 96u * g;
    const Half& hs = g ? h1 : h0;
    const uint32_t lw = o < 32u ? hs.l0 : (o < 64u ? hs.l1 : hs.l2);
    const uint32_t hw = o < 32u ? hs.h0 : (o < 64u ? hs.h1 : hs.h2);
    const uint32_t c = ((lw >> (o & 31u)) & 1u) | (((hw >> (o & 31u)) & 1u) << 1);
    const SymK k = sym_consts(c);
    uint32_t e0, l0, e1, l1;
    half_header(h0, c, 0u, e0, l0);
    half_header(h1, c, 1u, e1, l1);
    const uint32_t acc = half_counts(h0, off, k, 0u) + half_counts(h1, off, k, 1u);
    uint32_t rank = e0 + e1 + (acc & 0xFFu);
    if (c == 0u && r > primary) rank -= 1u;            // the '$' slot is stored as A
    return C[c] + rank;
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
