// Device-side FM-index construction: suffix array, BWT rank buckets of the text and of the
// reversed text, C / primary rows -- all on the GPU, so a 1 Gbp index takes seconds instead of minutes of host
// SA-IS (index_host.cpp).  Replaces ExactMatch.create_fm_index (reference SMEM/ExactMatch.py:22-33, 52-101:
// sort of all n rotations, n^2 memory) with
//   1. one LSD radix sort of (32-mer key, suffix start) pairs  -- cub::DeviceRadixSort, the only library call --
//   2. prefix doubling restricted to the rows whose key is not unique (Manber-Myers ranks, Larsson-Sadakane
//      style group refinement): group heads by max-scan, unresolved rows compacted, sorted by
//      (group head, rank[start + h]) and scattered back, h = 32, 64, ...; random DNA is done after step 1,
//      repeats cost log2(repeat length / 32) rounds over the repeated rows only;
//   3. BWT symbols gathered from the 2-bit text, ballot-packed into the bit planes of the 64-byte buckets
//      (fm_core.cuh layout) and checkpoint counts by a scan over per-bucket symbol counts.
// The result is bit-identical to the host builder (tests/test_gpu_parity.py compares SA and buckets).
// '$' handling: the text is 2-bit packed, so a suffix shorter than 32 bases is padded with A in its key and the
// tie with genuine A-runs is broken in the first doubling round by "virtual" ranks past the end of the text
// that decrease with the start position ('$' sorts below A, the shorter suffix first).
#include <cuda_runtime.h>

#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/genie_smem.h"
#include "fm_core.cuh"
#include "host_common.hpp"

namespace gsm {

#define GSM_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(GSM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

namespace {

constexpr int BT = 256;            // threads per block of the build kernels
constexpr int SUBTILES = 8;        // a tile = SUBTILES coalesced sub-tiles of BT items
constexpr int TILE = BT * SUBTILES;
constexpr uint32_t VIRT = 32;      // ranks are stored +VIRT; 0..VIRT-1 are the virtual ranks past the text end

struct MaxU {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

__device__ __forceinline__ uint32_t text_base(const uint32_t* words, uint64_t pos) {
    return (__ldg(words + (pos >> 4)) >> (30u - 2u * (uint32_t)(pos & 15u))) & 3u;
}

// ---------------------------------------------------------------------------------- text packing
// src: N bytes, ASCII ACGT (ascii != 0) or codes 0..3.  One thread per 16-base word; pad words are zeroed.
__global__ void __launch_bounds__(BT) k_pack_text(const uint8_t* src, uint64_t N, int ascii, uint32_t* words, uint64_t n_words,
                                                  unsigned long long* bad) {
    const uint64_t w = (uint64_t)blockIdx.x * BT + threadIdx.x;
    if (w >= n_words) return;
    uint32_t v = 0;
    const uint64_t base = w * 16;
    for (uint32_t t = 0; t < 16 && base + t < N; ++t) {
        uint32_t c = src[base + t];
        if (ascii) {
            const uint32_t ch = c;
            c = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
        }
        if (c > 3u) {
            atomicMin(bad, (unsigned long long)(base + t));
            c = 0;
        }
        v |= c << (30u - 2u * t);
    }
    words[w] = v;
}

// rtext = reverse(text), same packing.
__global__ void __launch_bounds__(BT) k_reverse_text(const uint32_t* text, uint64_t N, uint32_t* rtext, uint64_t n_words) {
    const uint64_t w = (uint64_t)blockIdx.x * BT + threadIdx.x;
    if (w >= n_words) return;
    uint32_t v = 0;
    const uint64_t base = w * 16;
    for (uint32_t t = 0; t < 16 && base + t < N; ++t) v |= text_base(text, N - 1 - (base + t)) << (30u - 2u * t);
    rtext[w] = v;
}

// ---------------------------------------------------------------------------------- sort keys
// key[i] = code of the 32 bases at suffix i (A-padded past the end), val[i] = i, for i in [0, N]; suffix N is '$'.
__global__ void __launch_bounds__(BT) k_suffix_keys(const uint32_t* text, uint64_t N, unsigned long long* keys, uint32_t* vals) {
    const uint64_t i = (uint64_t)blockIdx.x * BT + threadIdx.x;
    if (i > N) return;
    unsigned long long k = 0;
    if (i < N) k = kmer_code([&](uint64_t w) { return __ldg(text + w); }, i, 32);
    keys[i] = k;
    vals[i] = (uint32_t)i;
}

// ---------------------------------------------------------------------------------- group refinement
// Items k in [0, m): sorted keys, suffix starts vals[k], destination rows pos[k] (INIT: pos[k] = k).
// head(k) = row of the first item with the same key; an item is unresolved while its group has > 1 member.
template <bool INIT>
__device__ __forceinline__ void item_flags(const unsigned long long* keys, const uint32_t* pos, uint64_t k, uint64_t m, uint32_t& hv, uint32_t& unres) {
    const unsigned long long me = keys[k];
    const bool f = k == 0 || keys[k - 1] != me;
    const bool fn = k + 1 == m || keys[k + 1] != me;
    hv = f ? (INIT ? (uint32_t)k : pos[k]) : 0u;
    unres = (f && fn) ? 0u : 1u;
}

template <bool INIT>
__global__ void __launch_bounds__(BT) k_refine_reduce(const unsigned long long* keys, const uint32_t* pos, uint64_t m, uint32_t* tile_max,
                                                      uint32_t* tile_cnt) {
    using Reduce = cub::BlockReduce<uint32_t, BT>;
    __shared__ typename Reduce::TempStorage tmp_a, tmp_b;
    const uint64_t t0 = (uint64_t)blockIdx.x * TILE;
    uint32_t mx = 0, cnt = 0;
    for (int s = 0; s < SUBTILES; ++s) {
        const uint64_t k = t0 + (uint64_t)s * BT + threadIdx.x;
        if (k < m) {
            uint32_t hv, u;
            item_flags<INIT>(keys, pos, k, m, hv, u);
            mx = max(mx, hv);
            cnt += u;
        }
    }
    mx = Reduce(tmp_a).Reduce(mx, MaxU());
    cnt = Reduce(tmp_b).Sum(cnt);
    if (threadIdx.x == 0) {
        tile_max[blockIdx.x] = mx;
        tile_cnt[blockIdx.x] = cnt;
    }
}

// One block: exclusive max-prefix of tile_max (in place) and exclusive sum of tile_cnt -> tile_off; total -> *total.
__global__ void __launch_bounds__(1024) k_refine_top(uint32_t* tile_max, const uint32_t* tile_cnt, unsigned long long* tile_off, uint64_t n_tiles,
                                                     unsigned long long* total) {
    using ScanM = cub::BlockScan<uint32_t, 1024>;
    using ScanS = cub::BlockScan<unsigned long long, 1024>;
    __shared__ typename ScanM::TempStorage tm;
    __shared__ typename ScanS::TempStorage ts;
    uint32_t carry_m = 0;
    unsigned long long carry_s = 0;
    for (uint64_t base = 0; base < n_tiles; base += 1024) {
        const uint64_t t = base + threadIdx.x;
        const uint32_t vm = t < n_tiles ? tile_max[t] : 0u;
        const unsigned long long vs = t < n_tiles ? (unsigned long long)tile_cnt[t] : 0ull;
        uint32_t em, am;
        unsigned long long es, as;
        ScanM(tm).ExclusiveScan(vm, em, 0u, MaxU(), am);
        ScanS(ts).ExclusiveSum(vs, es, as);
        if (t < n_tiles) {
            tile_max[t] = max(em, carry_m);
            tile_off[t] = es + carry_s;
        }
        carry_m = max(carry_m, am);
        carry_s += as;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

// Final pass of a round: group heads by max-scan, rank[start] = head + VIRT, sa1[row] = start + 1 (the reference's
// 1-based values, ExactMatch.py:66), and the still unresolved items compacted into (head << 32, start, row).
template <bool INIT>
__global__ void __launch_bounds__(BT) k_refine_apply(const unsigned long long* keys, const uint32_t* vals, const uint32_t* pos, uint64_t m,
                                                     const uint32_t* tile_maxp, const unsigned long long* tile_off, uint32_t* rank, uint32_t* sa1,
                                                     unsigned long long* out_keys, uint32_t* out_vals, uint32_t* out_pos) {
    using ScanM = cub::BlockScan<uint32_t, BT>;
    using ScanS = cub::BlockScan<uint32_t, BT>;
    __shared__ typename ScanM::TempStorage tm;
    __shared__ typename ScanS::TempStorage ts;
    const uint64_t t0 = (uint64_t)blockIdx.x * TILE;
    uint32_t carry_m = tile_maxp[blockIdx.x];
    unsigned long long carry_s = tile_off[blockIdx.x];
    for (int s = 0; s < SUBTILES; ++s) {
        const uint64_t k = t0 + (uint64_t)s * BT + threadIdx.x;
        uint32_t hv = 0, u = 0;
        if (k < m) item_flags<INIT>(keys, pos, k, m, hv, u);
        uint32_t head, am, eo, as;
        ScanM(tm).InclusiveScan(hv, head, MaxU(), am);
        ScanS(ts).ExclusiveSum(u, eo, as);
        head = max(head, carry_m);
        if (k < m) {
            const uint32_t start = vals[k];
            const uint32_t row = INIT ? (uint32_t)k : pos[k];
            rank[start] = head + VIRT;
            sa1[row] = start + 1u;
            if (u) {
                const unsigned long long o = carry_s + eo;
                out_keys[o] = (unsigned long long)head << 32;
                out_vals[o] = start;
                out_pos[o] = row;
            }
        }
        carry_m = max(carry_m, am);
        carry_s += as;
        __syncthreads();
    }
}

// Low half of the doubling key: rank of the suffix h bases further on; past the text end the virtual rank
// VIRT - (p - N) (the '$' suffix N itself has a real rank).
__global__ void __launch_bounds__(BT) k_round_keys(unsigned long long* keys, const uint32_t* vals, const uint32_t* rank, uint64_t m, uint64_t h,
                                                   uint64_t N) {
    const uint64_t k = (uint64_t)blockIdx.x * BT + threadIdx.x;
    if (k >= m) return;
    const uint64_t p = (uint64_t)vals[k] + h;
    uint32_t r;
    if (p <= N) r = rank[p];
    else r = (p - N) <= VIRT ? VIRT - (uint32_t)(p - N) : 0u;
    keys[k] |= (unsigned long long)r;
}

// ---------------------------------------------------------------------------------- BWT -> rank buckets
// One warp per bucket (192 rows): 6 coalesced SA reads, one text gather per row, ballots form the bit planes.
__global__ void __launch_bounds__(BT) k_bwt_planes(const uint32_t* sa1, const uint32_t* text, uint64_t n_rows, uint64_t n_buckets, uint32_t* buckets,
                                                   uint4* bucket_cnt, uint32_t* primary) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * BT + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * BT) >> 5;
    for (uint64_t b = warp0; b < n_buckets; b += n_warps) {
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        uint32_t* w = buckets + b * 16;
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
            for (int mth = 0; mth < 3; ++mth) {
                const uint64_t r = b * GSM_BUCKET_SYMS + 96 * g + 32 * mth + lane;
                const bool valid = r < n_rows;
                uint32_t c = 0;
                if (valid) {
                    const uint32_t s1 = __ldg(sa1 + r);           // 1-based start; the suffix starting at base 0 has '$' before it
                    if (s1 == 1u) *primary = (uint32_t)r;
                    else c = text_base(text, (uint64_t)s1 - 2u);
                }
                const uint32_t lo = __ballot_sync(0xFFFFFFFFu, valid && (c & 1u));
                const uint32_t hi = __ballot_sync(0xFFFFFFFFu, valid && (c >> 1));
                const uint32_t vm = __ballot_sync(0xFFFFFFFFu, valid);
                const uint32_t n3 = __popc(lo & hi), n2 = __popc(hi & ~lo), n1 = __popc(lo & ~hi);
                c3 += n3; c2 += n2; c1 += n1; c0 += __popc(vm) - n1 - n2 - n3;
                if (lane == 0) {
                    w[8 * g + 2 + mth] = lo;
                    w[8 * g + 5 + mth] = hi;
                }
            }
        if (lane == 0) bucket_cnt[b] = make_uint4(c0, c1, c2, c3);
    }
}

struct Add4 {
    __device__ __forceinline__ uint4 operator()(const uint4& a, const uint4& b) const { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
};

__global__ void __launch_bounds__(BT) k_cnt_reduce(const uint4* cnt, uint64_t n, uint4* tile_sum) {
    using Reduce = cub::BlockReduce<uint4, BT>;
    __shared__ typename Reduce::TempStorage tmp;
    const uint64_t t0 = (uint64_t)blockIdx.x * TILE;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < SUBTILES; ++s) {
        const uint64_t k = t0 + (uint64_t)s * BT + threadIdx.x;
        if (k < n) acc = Add4()(acc, cnt[k]);
    }
    acc = Reduce(tmp).Reduce(acc, Add4());
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(1024) k_cnt_top(uint4* tile_sum, uint64_t n_tiles, uint4* total) {
    using Scan = cub::BlockScan<uint4, 1024>;
    __shared__ typename Scan::TempStorage tmp;
    uint4 carry = make_uint4(0, 0, 0, 0);
    for (uint64_t base = 0; base < n_tiles; base += 1024) {
        const uint64_t t = base + threadIdx.x;
        const uint4 v = t < n_tiles ? tile_sum[t] : make_uint4(0, 0, 0, 0);
        uint4 ex, agg;
        Scan(tmp).ExclusiveScan(v, ex, make_uint4(0, 0, 0, 0), Add4(), agg);
        if (t < n_tiles) tile_sum[t] = Add4()(ex, carry);
        carry = Add4()(carry, agg);
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// checkpoint words of bucket b = symbol counts of rows [0, 192 b)  ('$' slot counted as A, index_host.cpp pack_buckets)
__global__ void __launch_bounds__(BT) k_cnt_apply(const uint4* cnt, uint64_t n, const uint4* tile_pre, uint32_t* buckets) {
    using Scan = cub::BlockScan<uint4, BT>;
    __shared__ typename Scan::TempStorage tmp;
    const uint64_t t0 = (uint64_t)blockIdx.x * TILE;
    uint4 carry = tile_pre[blockIdx.x];
    for (int s = 0; s < SUBTILES; ++s) {
        const uint64_t k = t0 + (uint64_t)s * BT + threadIdx.x;
        const uint4 v = k < n ? cnt[k] : make_uint4(0, 0, 0, 0);
        uint4 ex, agg;
        Scan(tmp).ExclusiveScan(v, ex, make_uint4(0, 0, 0, 0), Add4(), agg);
        if (k < n) {
            const uint4 p = Add4()(ex, carry);
            uint32_t* w = buckets + k * 16;
            w[0] = p.x; w[1] = p.y; w[8] = p.z; w[9] = p.w;
        }
        carry = Add4()(carry, agg);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- workspace carving
struct Carve {
    uint8_t* base;
    uint64_t off = 0;
    explicit Carve(void* b) : base((uint8_t*)b) {}
    template <typename T>
    T* take(uint64_t count) {
        off = (off + 255) & ~255ull;
        T* p = base ? (T*)(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct Plan {
    unsigned long long *k0, *k1, *tile_off, *scalars;
    uint32_t *v0, *v1, *p0, *p1, *rank, *rtext, *tile_max, *tile_cnt;
    uint4 *bucket_cnt, *cnt_tiles;
    void* cub_tmp;
    uint64_t cub_bytes, total;
};

uint64_t cub_sort_bytes(uint64_t n) {
    size_t bytes = 0;
    cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, (uint64_t)n, 0, 64, (cudaStream_t)0);
    return (uint64_t)bytes;
}

Plan make_plan(void* ws, uint64_t N, uint32_t flags) {
    const uint64_t n = N + 1;
    const uint64_t n_tiles = (n + TILE - 1) / TILE;
    const uint64_t nb = n / GSM_BUCKET_SYMS + 1;
    const uint64_t text_words = (N + 15) / 16 + 2;
    Carve c(ws);
    Plan p;
    p.k0 = c.take<unsigned long long>(n);
    p.k1 = c.take<unsigned long long>(n);
    p.v0 = c.take<uint32_t>(n);
    p.v1 = c.take<uint32_t>(n);
    p.p0 = c.take<uint32_t>(n);
    p.p1 = c.take<uint32_t>(n);
    p.rank = c.take<uint32_t>(n);
    p.rtext = c.take<uint32_t>((flags & 1u) ? text_words : 0);
    p.tile_max = c.take<uint32_t>(n_tiles);
    p.tile_cnt = c.take<uint32_t>(n_tiles);
    p.tile_off = c.take<unsigned long long>(n_tiles);
    p.bucket_cnt = c.take<uint4>(nb);
    p.cnt_tiles = c.take<uint4>((nb + TILE - 1) / TILE);
    p.scalars = c.take<unsigned long long>(16);
    p.cub_bytes = cub_sort_bytes(n);
    p.cub_tmp = c.take<uint8_t>(p.cub_bytes);
    p.total = (c.off + 255) & ~255ull;
    return p;
}

inline unsigned grid_for(uint64_t items, uint64_t per_block) { return (unsigned)((items + per_block - 1) / per_block); }

// Suffix array of `text` (N bases + '$') into sa1 (1-based starts, n rows).  rounds_out: doubling rounds used.
int suffix_array_device(const uint32_t* text, uint64_t N, uint32_t* sa1, const Plan& p, cudaStream_t st, uint32_t* rounds_out) {
    const uint64_t n = N + 1;
    k_suffix_keys<<<grid_for(n, BT), BT, 0, st>>>(text, N, p.k0, p.v0);
    cub::DoubleBuffer<unsigned long long> dk(p.k0, p.k1);
    cub::DoubleBuffer<uint32_t> dv(p.v0, p.v1);
    size_t tmp_bytes = p.cub_bytes;
    GSM_CUDA(cub::DeviceRadixSort::SortPairs(p.cub_tmp, tmp_bytes, dk, dv, (uint64_t)n, 0, 64, st));
    uint32_t* pos_in = p.p0;
    uint32_t* pos_out = p.p1;
    unsigned long long* total_dev = p.scalars;
    uint64_t m = n;
    uint64_t h = 32;
    uint32_t rounds = 0;
    bool init = true;
    int key_bits = 33;
    while ((1ull << (key_bits - 32)) < n + VIRT && key_bits < 64) ++key_bits;      // bits of the doubling keys that can be set
    while (true) {
        const uint64_t n_tiles = (m + TILE - 1) / TILE;
        unsigned long long* keys = dk.Current();
        uint32_t* vals = dv.Current();
        unsigned long long* okeys = dk.Alternate();
        uint32_t* ovals = dv.Alternate();
        if (init) {
            k_refine_reduce<true><<<(unsigned)n_tiles, BT, 0, st>>>(keys, nullptr, m, p.tile_max, p.tile_cnt);
            k_refine_top<<<1, 1024, 0, st>>>(p.tile_max, p.tile_cnt, p.tile_off, n_tiles, total_dev);
            k_refine_apply<true><<<(unsigned)n_tiles, BT, 0, st>>>(keys, vals, nullptr, m, p.tile_max, p.tile_off, p.rank, sa1, okeys, ovals, pos_out);
        } else {
            k_refine_reduce<false><<<(unsigned)n_tiles, BT, 0, st>>>(keys, pos_in, m, p.tile_max, p.tile_cnt);
            k_refine_top<<<1, 1024, 0, st>>>(p.tile_max, p.tile_cnt, p.tile_off, n_tiles, total_dev);
            k_refine_apply<false><<<(unsigned)n_tiles, BT, 0, st>>>(keys, vals, pos_in, m, p.tile_max, p.tile_off, p.rank, sa1, okeys, ovals, pos_out);
        }
        unsigned long long total = 0;
        GSM_CUDA(cudaMemcpyAsync(&total, total_dev, sizeof(total), cudaMemcpyDeviceToHost, st));
        GSM_CUDA(cudaStreamSynchronize(st));
        m = total;
        if (m == 0) break;
        if (h > 2 * n + 64) return fail(GSM_E_CUDA, "gsm_index_build_device: prefix doubling did not converge (internal error)");
        init = false;
        dk = cub::DoubleBuffer<unsigned long long>(okeys, keys);
        dv = cub::DoubleBuffer<uint32_t>(ovals, vals);
        { uint32_t* t = pos_in; pos_in = pos_out; pos_out = t; }
        k_round_keys<<<grid_for(m, BT), BT, 0, st>>>(dk.Current(), dv.Current(), p.rank, m, h, N);
        tmp_bytes = p.cub_bytes;
        GSM_CUDA(cub::DeviceRadixSort::SortPairs(p.cub_tmp, tmp_bytes, dk, dv, (uint64_t)m, 0, key_bits, st));
        h *= 2;
        ++rounds;
    }
    GSM_CUDA(cudaGetLastError());
    if (rounds_out) *rounds_out = rounds;
    return GSM_OK;
}

// sa1 + text -> packed buckets; totals (4 x u32, '$' counted as A) -> scalars[2..3], primary -> scalars[4] (as u32)
int buckets_device(const uint32_t* sa1, const uint32_t* text, uint64_t N, uint32_t* buckets, const Plan& p, cudaStream_t st, uint4* totals_dev,
                   uint32_t* primary_dev) {
    const uint64_t n = N + 1;
    const uint64_t nb = n / GSM_BUCKET_SYMS + 1;
    const uint64_t nt = (nb + TILE - 1) / TILE;
    const unsigned grid = (unsigned)std::min<uint64_t>((nb + (BT / 32) - 1) / (BT / 32), 148ull * 64);
    k_bwt_planes<<<grid, BT, 0, st>>>(sa1, text, n, nb, buckets, p.bucket_cnt, primary_dev);
    k_cnt_reduce<<<(unsigned)nt, BT, 0, st>>>(p.bucket_cnt, nb, p.cnt_tiles);
    k_cnt_top<<<1, 1024, 0, st>>>(p.cnt_tiles, nt, totals_dev);
    k_cnt_apply<<<(unsigned)nt, BT, 0, st>>>(p.bucket_cnt, nb, p.cnt_tiles, buckets);
    GSM_CUDA(cudaGetLastError());
    return GSM_OK;
}

}  // namespace
}  // namespace gsm

using namespace gsm;

extern "C" {

int gsm_text_pack_device(const void* bases_dev, uint64_t n_bases, uint32_t ascii, uint32_t* text2bit, uint64_t* scratch8, void* stream) {
    if (!bases_dev || !text2bit || !scratch8 || n_bases == 0) return fail(GSM_E_INVALID, "gsm_text_pack_device: null/empty input");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t words = (n_bases + 15) / 16 + 2;
    GSM_CUDA(cudaMemsetAsync(scratch8, 0xFF, 8, st));
    k_pack_text<<<grid_for(words, BT), BT, 0, st>>>((const uint8_t*)bases_dev, n_bases, (int)ascii, text2bit, words, (unsigned long long*)scratch8);
    GSM_CUDA(cudaGetLastError());
    unsigned long long bad = 0;
    GSM_CUDA(cudaMemcpyAsync(&bad, scratch8, 8, cudaMemcpyDeviceToHost, st));
    GSM_CUDA(cudaStreamSynchronize(st));
    if (bad != ~0ull) return fail(GSM_E_INVALID, "non-ACGT base at offset " + std::to_string(bad));
    return GSM_OK;
}

int gsm_index_build_device_workspace(uint64_t n_bases, uint32_t flags, uint64_t* bytes) {
    if (!bytes || n_bases == 0) return fail(GSM_E_INVALID, "gsm_index_build_device_workspace: null/empty input");
    if (n_bases + 1 + VIRT >= (1ull << 32)) return fail(GSM_E_INVALID, "n_bases must be < 2^32 - 34 (32-bit rows)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GSM_E_NODEVICE, "no CUDA device: the index builder has no CPU fallback here (use gsm_index_build)");
    *bytes = make_plan(nullptr, n_bases, flags).total;
    return GSM_OK;
}

int gsm_index_build_device(const uint32_t* text2bit, uint64_t n_bases, uint32_t flags, uint32_t* sa, void* fwd_buckets, void* rev_buckets,
                           void* workspace, uint64_t workspace_bytes, gsm_index_info* info, void* stream) {
    if (!text2bit || !fwd_buckets || !workspace || !info || n_bases == 0) return fail(GSM_E_INVALID, "gsm_index_build_device: null/empty input");
    if (n_bases + 1 + VIRT >= (1ull << 32)) return fail(GSM_E_INVALID, "n_bases must be < 2^32 - 34 (32-bit rows)");
    const bool want_rev = (flags & 1u) != 0;
    if (want_rev && !rev_buckets) return fail(GSM_E_INVALID, "gsm_index_build_device: rev_buckets is NULL but flags bit0 asks for the reverse index");
    if (!sa) return fail(GSM_E_INVALID, "gsm_index_build_device: sa is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(GSM_E_NODEVICE, "no CUDA device");
    const Plan p = make_plan(workspace, n_bases, flags);
    if (workspace_bytes < p.total) return fail(GSM_E_CAPACITY, "gsm_index_build_device: workspace too small, need " + std::to_string(p.total) + " bytes");
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t N = n_bases, n = N + 1;
    uint4* totals_dev = (uint4*)(p.scalars + 2);
    uint32_t* prim_dev = (uint32_t*)(p.scalars + 4);
    uint32_t rounds = 0;
    memset(info, 0, sizeof(*info));
    uint32_t prim[2] = {0, 0};
    // the reverse index first: its suffix array is scratch (it lives in the caller's sa buffer until the forward one overwrites it)
    if (want_rev) {
        uint32_t* rsa = sa;
        k_reverse_text<<<grid_for((N + 15) / 16 + 2, BT), BT, 0, st>>>(text2bit, N, p.rtext, (N + 15) / 16 + 2);
        int s = suffix_array_device(p.rtext, N, rsa, p, st, &rounds);
        if (s != GSM_OK) return s;
        s = buckets_device(rsa, p.rtext, N, (uint32_t*)rev_buckets, p, st, totals_dev, prim_dev + 1);
        if (s != GSM_OK) return s;
    }
    uint32_t* fsa = sa;
    int s = suffix_array_device(text2bit, N, fsa, p, st, &rounds);
    if (s != GSM_OK) return s;
    s = buckets_device(fsa, text2bit, N, (uint32_t*)fwd_buckets, p, st, totals_dev, prim_dev);
    if (s != GSM_OK) return s;
    uint32_t tot[4];
    GSM_CUDA(cudaMemcpyAsync(tot, totals_dev, sizeof(tot), cudaMemcpyDeviceToHost, st));
    GSM_CUDA(cudaMemcpyAsync(prim, prim_dev, sizeof(prim), cudaMemcpyDeviceToHost, st));
    GSM_CUDA(cudaStreamSynchronize(st));
    info->n_bases = N;
    info->n_rows = n;
    info->n_buckets = n / GSM_BUCKET_SYMS + 1;
    info->bucket_bytes = info->n_buckets * GSM_BUCKET_BYTES;
    info->text_words = (N + 15) / 16 + 2;
    tot[0] -= 1;                                   // the '$' slot was counted as A
    for (int c = 0; c < 4; ++c) info->count[c] = tot[c];
    info->C[0] = 1;
    for (int c = 1; c <= 4; ++c) info->C[c] = info->C[c - 1] + info->count[c - 1];
    info->primary_fwd = prim[0];
    info->primary_rev = prim[1];
    info->has_reverse = want_rev ? 1 : 0;
    info->reserved = rounds;
    return GSM_OK;
}

}  // extern "C"
