// Device side of the sweep kernel.  Two lanes ("a pair") cooperate on every 64-byte rank bucket
// of one read: each issues ONE 256-bit load for its half, popcounts up to 96 symbols and the two
// halves are summed with one shuffle.  16 reads are in flight per warp, and the whole warp executes
// ONE uniform load/popcount section per iteration (pair_step) so that all 16 chains' bucket fetches
// are in flight together; the per-read control (sweep_logic.cuh) only touches registers and shared
// memory.  (A 4-lane variant with 128-bit loads was measured first: 78 M reads/s against 103 M for
// pairs on the same bench, because the kernel is instruction-issue bound, not bandwidth bound --
// profiles/r01_notes.md.)
#pragma once
#include <cuda_runtime.h>

#include "sweep_logic.cuh"

namespace gsm {

constexpr int SWEEP_THREADS = 128;
constexpr int SWEEP_LPR = 2;                              // lanes per read
constexpr int SWEEP_GROUPS = SWEEP_THREADS / SWEEP_LPR;   // reads in flight per block
constexpr int SWEEP_CAP = 12;                             // candidates kept in shared memory per read
constexpr int SWEEP_MIN_BLOCKS = 8;

struct SweepArgs {
    const uint4* fwd;
    const uint4* rev;
    IndexMeta meta;
    const uint4* reads;
    const uint32_t* chunk_off;
    const uint32_t* len;
    uint32_t n_reads;
    uint32_t read_u4;        // uint4 slots of shared memory per read for its unpacked bases
    uint32_t pack_u4;        // uint4 slots per read for its packed words (k-mer codes of the seed table)
    uint32_t max_len;
    const uint4* seed_tab;   // optional 4^seed_K x {fwd lo, count, rev lo, 0} (gsm_seed_table_build), else NULL
    uint32_t seed_K;
    uint4* mem_pool;
    unsigned long long mem_cap;
    uint32_t* mem_off;
    uint32_t* mem_cnt;
    uint4* scratch;          // per pair: [0, max_len) match staging, [max_len, 2 max_len) candidate spill
    unsigned long long* counters;
};

// Shared memory: [per pair: SWEEP_CAP candidates of 16 B {end, lo, cnt, -}, then
// the read's bases one per byte, then its packed words][16 B pad].  Indexed through one extern array so that
// the compiler emits LDS/STS.
extern __shared__ uint4 g_sweep_smem[];

// one 32-byte half of a bucket: a single 256-bit read-only load
__device__ __forceinline__ Half ldg_half(const uint4* halves_base, size_t half_index) {
    Half h;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(h.c0), "=r"(h.c1), "=r"(h.l0), "=r"(h.l1), "=r"(h.l2), "=r"(h.h0), "=r"(h.h1), "=r"(h.h2)
                 : "l"(halves_base + half_index * 2));
    return h;
}

// One FM extension step executed by all pairs of a warp together.  Every lane passes its pair's
// operands; `active` pairs get their result, inactive ones issue no loads.  g = lane within pair.
__device__ __forceinline__ StepOut pair_step(const uint4* __restrict__ bk, uint32_t P0, uint32_t P1, uint32_t ch, uint32_t Cc,
                                             uint32_t primary, uint32_t g, bool active) {
    constexpr uint32_t FULLM = 0xFFFFFFFFu;
    uint32_t b0, r0, b1, r1;
    split192(P0, b0, r0);
    split192(P1, b1, r1);
    Half v0 = Half{0, 0, 0, 0, 0, 0, 0, 0}, v1;
    if (active) v0 = ldg_half(bk, (size_t)b0 * 2 + g);
    v1 = v0;
    if (active && b1 != b0) v1 = ldg_half(bk, (size_t)b1 * 2 + g);
    const SymK k = sym_consts(ch);
    uint32_t packed = half_counts(v0, r0, k, g) | (half_counts(v1, r1, k, g) << 16);
    uint32_t he0, hl0, he1, hl1;
    half_header(v0, ch, g, he0, hl0);
    half_header(v1, ch, g, he1, hl1);
    uint32_t A = he0, B = he1, D = hl1 - hl0;
    packed += __shfl_xor_sync(FULLM, packed, 1);
    A += __shfl_xor_sync(FULLM, A, 1);
    B += __shfl_xor_sync(FULLM, B, 1);
    D += __shfl_xor_sync(FULLM, D, 1);
    const uint32_t eq0 = A + (packed & 0xFFu);
    const uint32_t eq1 = B + ((packed >> 16) & 0xFFu);
    const uint32_t ltd = D + ((packed >> 24) & 0xFFu) - ((packed >> 8) & 0xFFu);
    return finish_step(eq0, eq1, ltd, P0, P1, ch, Cc, primary);
}

// one 16-byte seed-table entry (read-only path; volatile so that it is issued before the bucket loads' consumers)
__device__ __forceinline__ uint4 ldg_seed(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// byte of four 2-bit bases (MSB first) -> four bytes, first base at the lowest address (no table: one multiply)
__device__ __forceinline__ uint32_t spread4(uint32_t x) { return ((x * 0x01004010u) | (x >> 6)) & 0x03030303u; }

struct DevSweepCtx {
    const SweepArgs& a;
    uint32_t cand0;       // index (uint4) of this pair's candidate slots
    uint32_t bytes0;      // byte offset of this pair's unpacked read
    uint32_t words0;      // word offset of this pair's packed read
    uint4* stage;         // global: match staging of this pair
    uint4* spill;         // global: candidate spill of this pair
    uint32_t g, gmask, gbase;

    __device__ __forceinline__ bool fetch(uint32_t& rid, uint32_t& L) {
        unsigned long long r = 0;
        if (g == 0) r = atomicAdd(&a.counters[3], 1ull);
        r = __shfl_sync(gmask, r, gbase);
        if (r >= a.n_reads) return false;
        rid = (uint32_t)r;
        L = __ldg(a.len + rid);
        const uint32_t off = __ldg(a.chunk_off + rid);
        const uint32_t nch = (L + 63u) >> 6;
        uint4* dst = g_sweep_smem + (bytes0 >> 4);
        __syncwarp(gmask);                       // both lanes are done with the previous read's bases
        for (uint32_t c = g; c < nch; c += SWEEP_LPR) {
            const uint4 v = __ldg(a.reads + (size_t)off + c);
            g_sweep_smem[(words0 >> 2) + c] = v;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)          // 16 packed bases -> 16 bytes, first base at the lowest address
                if (c * 4 + k < a.read_u4)
                    dst[c * 4 + k] = make_uint4(spread4(w[k] >> 24), spread4((w[k] >> 16) & 0xFFu), spread4((w[k] >> 8) & 0xFFu), spread4(w[k] & 0xFFu));
        }
        __syncwarp(gmask);
        return true;
    }
    __device__ __forceinline__ uint32_t base(uint32_t pos) const {
        return reinterpret_cast<const uint8_t*>(g_sweep_smem)[bytes0 + pos];
    }
    __device__ __forceinline__ uint32_t seed_k() const { return a.seed_K; }
    // the unique-match shortcut of sweep_logic.cuh stays compiled out of the kernel: measured slower (profiles/r01_notes.md)
    __device__ __forceinline__ constexpr bool uniq() const { return false; }
    // code of q[pos:pos+K): top 2K bits of the 64-bit window starting at base pos (MSB-first packing)
    __device__ __forceinline__ uint32_t kmer(uint32_t pos) const {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(g_sweep_smem) + words0 + (pos >> 4);
        return __funnelshift_l(w[1], w[0], 2u * (pos & 15u)) >> (32u - 2u * a.seed_K);
    }
    __device__ __forceinline__ void cand_put(uint32_t i, uint32_t j, uint32_t lo, uint32_t cnt) {
        if (i < (uint32_t)SWEEP_CAP) g_sweep_smem[cand0 + i] = make_uint4(j, lo, cnt, 0u);   // same value from both lanes
        else if (g == 0) __stcg(spill + (i - SWEEP_CAP), make_uint4(j, lo, cnt, 0u));
    }
    __device__ __forceinline__ void cand_get(uint32_t i, uint32_t& j, uint32_t& lo, uint32_t& cnt) const {
        uint4 v;
        if (i < (uint32_t)SWEEP_CAP) v = g_sweep_smem[cand0 + i];
        else v = __ldcg(spill + (i - SWEEP_CAP));
        j = v.x; lo = v.y; cnt = v.z;
    }
    __device__ __forceinline__ void cand_sync() { __syncwarp(gmask); }
    __device__ __forceinline__ void emit(uint32_t idx, MemEntry e) {
        if (g == 0) __stcg(stage + idx, make_uint4(e.se, e.lo, e.cnt, e.sweep));
    }
    __device__ __forceinline__ void finish(uint32_t rid, uint32_t n) {
        __syncwarp(gmask);
        unsigned long long off = 0;
        if (g == 0) off = atomicAdd(&a.counters[0], (unsigned long long)n);
        off = __shfl_sync(gmask, off, gbase);
        if (off + n > a.mem_cap) {
            if (g == 0) { atomicOr(&a.counters[2], 1ull); a.mem_off[rid] = 0; a.mem_cnt[rid] = 0; }
            return;
        }
        for (uint32_t k = g; k < n; k += SWEEP_LPR) a.mem_pool[off + k] = __ldcg(stage + k);
        if (g == 0) { a.mem_off[rid] = (uint32_t)off; a.mem_cnt[rid] = n; }
    }
};

// one byte per base, rounded up to 16 bytes
__host__ __device__ inline uint32_t sweep_read_u4(uint32_t max_len) { return (max_len + 15u) / 16u; }

__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_BLOCKS) k_sweep(const SweepArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t g = lane & 1u;
    const uint32_t pair_in_block = threadIdx.x >> 1;
    const uint32_t pair_u4 = SWEEP_CAP + a.read_u4 + a.pack_u4;
    const uint32_t p0 = pair_in_block * pair_u4;
    const size_t gp = (size_t)blockIdx.x * SWEEP_GROUPS + pair_in_block;
    DevSweepCtx ctx{a, p0, (p0 + SWEEP_CAP) * 16u, (p0 + SWEEP_CAP + a.read_u4) * 4u, a.scratch + gp * 2 * a.max_len,
                    a.scratch + gp * 2 * a.max_len + a.max_len, g, 3u << (lane & ~1u), lane & ~1u};
    Sweeper<DevSweepCtx> sw;
    for (;;) {
        const bool need = sw.next(ctx, a.meta);
        if (!__any_sync(0xFFFFFFFFu, need)) break;
        // ONE uniform memory section per iteration: every pair issues its pending fetch here -- a seed-table entry
        // or the (at most two) buckets of an FM step -- so all 16 chains' loads are in flight together.
        const bool is_seed = need && sw.pending_seed();
        const bool is_step = need && !is_seed;
        uint4 se = make_uint4(0u, 0u, 0u, 0u);
        if (is_seed) se = ldg_seed(a.seed_tab + sw.P0);
        const bool rev = sw.on_reverse();
        const StepOut r = pair_step(rev ? a.rev : a.fwd, sw.P0, sw.P0 + sw.cnt, sw.ch, a.meta.C[sw.ch & 3u],
                                    rev ? a.meta.prim_r : a.meta.prim_f, g, is_step);
        if (is_step) sw.consume(ctx, a.meta, r);
        else if (is_seed) sw.consume_seed(ctx, a.meta, SeedEntry{se.x, se.y, se.z, se.w});
    }
}

// packed words of a read, rounded up to 16-byte chunks
__host__ __device__ inline uint32_t sweep_pack_u4(uint32_t max_len) { return (max_len + 63u) / 64u; }

inline size_t sweep_smem_bytes(uint32_t max_len) {
    return (size_t)SWEEP_GROUPS * (SWEEP_CAP + sweep_read_u4(max_len) + sweep_pack_u4(max_len)) * sizeof(uint4) + 16;
}

}  // namespace gsm
