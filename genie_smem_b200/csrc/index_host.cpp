// Host-side FM-index construction for the B200 SMEM engine.
//
// Replaces ExactMatch.create_fm_index (reference SMEM/ExactMatch.py:22-33): the reference sorts
// all n rotations as n strings of length n (ExactMatch.py:52-58, n^2 memory); here the suffix
// array of text+'$' comes from an O(n) induced-sorting construction (SA-IS, Nong/Zhang/Chan),
// so 10^8..10^9-base references are reachable.  The arrays it yields are the reference's:
//   suffix_array (1-based starts, SA[0] = n)      ExactMatch.py:66
//   bwt_array    (last column)                    ExactMatch.py:64
//   count_dic    (first row per leading char)     ExactMatch.py:92-101
// The inclusive occurrence matrix (ExactMatch.py:70-90) is not materialised: it is replaced by
// 64-byte rank buckets (two 32-byte halves of {u32 occ[2]; 96 two-bit symbols as bit planes}, fm_core.cuh).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/genie_smem.h"
#include "host_common.hpp"

namespace gsm {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// ---------------------------------------------------------------------------------- SA-IS
// s: n symbols in [0, K), s[n-1] is the unique smallest symbol (the sentinel).  SA: n entries.
namespace {

template <typename Ch>
struct Sais {
    const Ch* s;
    int32_t* SA;
    int32_t n;
    int32_t K;
    std::vector<uint64_t> tbits;  // 1 = S-type
    std::vector<int32_t> bkt;

    inline bool tget(int32_t i) const { return (tbits[(uint32_t)i >> 6] >> (i & 63)) & 1u; }
    inline void tset(int32_t i, bool b) {
        uint64_t m = 1ull << (i & 63);
        if (b) tbits[(uint32_t)i >> 6] |= m; else tbits[(uint32_t)i >> 6] &= ~m;
    }
    inline bool is_lms(int32_t i) const { return i > 0 && tget(i) && !tget(i - 1); }

    void buckets(bool end) {
        std::fill(bkt.begin(), bkt.end(), 0);
        for (int32_t i = 0; i < n; ++i) bkt[s[i]]++;
        int32_t sum = 0;
        for (int32_t c = 0; c < K; ++c) {
            sum += bkt[c];
            bkt[c] = end ? sum : sum - bkt[c];
        }
    }
    void induce_l() {
        buckets(false);
        for (int32_t i = 0; i < n; ++i) {
            int32_t j = SA[i] - 1;
            if (j >= 0 && !tget(j)) SA[bkt[s[j]]++] = j;
        }
    }
    void induce_s() {
        buckets(true);
        for (int32_t i = n - 1; i >= 0; --i) {
            int32_t j = SA[i] - 1;
            if (j >= 0 && tget(j)) SA[--bkt[s[j]]] = j;
        }
    }

    void run() {
        tbits.assign(((size_t)n + 63) / 64, 0);
        bkt.assign(K, 0);
        if (n == 1) { SA[0] = 0; return; }
        tset(n - 1, true);
        tset(n - 2, false);
        for (int32_t i = n - 3; i >= 0; --i)
            tset(i, s[i] < s[i + 1] || (s[i] == s[i + 1] && tget(i + 1)));

        // stage 1: sort the LMS substrings by one round of induced sorting
        buckets(true);
        std::fill(SA, SA + n, -1);
        for (int32_t i = 1; i < n; ++i)
            if (is_lms(i)) SA[--bkt[s[i]]] = i;
        induce_l();
        induce_s();

        int32_t n1 = 0;
        for (int32_t i = 0; i < n; ++i)
            if (is_lms(SA[i])) SA[n1++] = SA[i];
        std::fill(SA + n1, SA + n, -1);
        int32_t name = 0, prev = -1;
        for (int32_t i = 0; i < n1; ++i) {
            int32_t pos = SA[i];
            bool diff = false;
            for (int32_t d = 0; d < n; ++d) {
                if (prev == -1 || s[pos + d] != s[prev + d] || tget(pos + d) != tget(prev + d)) { diff = true; break; }
                if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
            }
            if (diff) { ++name; prev = pos; }
            SA[n1 + pos / 2] = name - 1;
        }
        for (int32_t i = n - 1, j = n - 1; i >= n1; --i)
            if (SA[i] >= 0) SA[j--] = SA[i];

        // stage 2: order the LMS suffixes (recurse on the reduced string if names collide)
        int32_t* SA1 = SA;
        int32_t* s1 = SA + n - n1;
        if (name < n1) {
            Sais<int32_t> sub;
            sub.s = s1; sub.SA = SA1; sub.n = n1; sub.K = name;
            sub.run();
        } else {
            for (int32_t i = 0; i < n1; ++i) SA1[s1[i]] = i;
        }

        // stage 3: induce the full order from the sorted LMS suffixes
        buckets(true);
        for (int32_t i = 1, j = 0; i < n; ++i)
            if (is_lms(i)) s1[j++] = i;
        for (int32_t i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
        std::fill(SA + n1, SA + n, -1);
        for (int32_t i = n1 - 1; i >= 0; --i) {
            int32_t j = SA[i];
            SA[i] = -1;
            SA[--bkt[s[j]]] = j;
        }
        induce_l();
        induce_s();
    }
};

}  // namespace

void suffix_sort(const uint8_t* sym, int32_t* SA, int32_t n) {
    Sais<uint8_t> top;
    top.s = sym; top.SA = SA; top.n = n; top.K = 5;
    top.run();
}

// ---------------------------------------------------------------------------------- buckets
// bwt: n symbols in 0..3 ('$' stored as 0 at the primary row).  Layout per 64-byte bucket b (rows
// [192b, 192b+192)), two 32-byte halves g = 0, 1 (fm_core.cuh): w[8g], w[8g+1] = raw counts of symbols
// 2g, 2g+1 in rows [0,192b) (the '$' slot counts as A; rank(A, p) is corrected by (p > primary) on the
// device); w[8g+2..8g+4] = low-bit plane of symbols 96g..96g+95, w[8g+5..8g+7] = high-bit plane.
void pack_buckets(const uint8_t* bwt, uint64_t n, std::vector<uint32_t>& out) {
    uint64_t nb = n / GSM_BUCKET_SYMS + 1;
    out.assign(nb * 16, 0);
    uint32_t run[4] = {0, 0, 0, 0};
    for (uint64_t b = 0; b < nb; ++b) {
        uint32_t* w = &out[b * 16];
        w[0] = run[0]; w[1] = run[1]; w[8] = run[2]; w[9] = run[3];
        uint64_t base = b * GSM_BUCKET_SYMS;
        for (int g = 0; g < 2; ++g)
            for (int m = 0; m < 3; ++m) {
                uint32_t lo = 0, hi = 0;
                for (int t = 0; t < 32; ++t) {
                    uint64_t r = base + 96 * g + 32 * m + t;
                    if (r >= n) break;
                    uint32_t c = bwt[r];
                    run[c]++;
                    lo |= (c & 1u) << t;
                    hi |= (c >> 1) << t;
                }
                w[8 * g + 2 + m] = lo;
                w[8 * g + 5 + m] = hi;
            }
    }
}

}  // namespace gsm

using namespace gsm;

struct gsm_index {
    uint64_t n_bases = 0;
    uint64_t n = 0;                 // rows
    std::vector<uint8_t> codes;     // n_bases symbols 0..3
    std::vector<int32_t> sa0;       // 0-based suffix starts, n entries
    std::vector<uint32_t> fwd, rev; // packed buckets
    uint32_t count[4] = {0, 0, 0, 0};
    uint32_t C[5] = {0, 0, 0, 0, 0};
    uint32_t primary_fwd = 0, primary_rev = 0;
    bool has_rev = false;
};

static int decode_bases(const char* bases, uint64_t n_bases, std::vector<uint8_t>& codes, uint32_t count[4]) {
    static int8_t map[256];
    static bool init = false;
    if (!init) {
        memset(map, -1, sizeof(map));
        map[(int)'A'] = 0; map[(int)'C'] = 1; map[(int)'G'] = 2; map[(int)'T'] = 3;
        init = true;
    }
    codes.resize(n_bases);
    for (uint64_t i = 0; i < n_bases; ++i) {
        int8_t c = map[(uint8_t)bases[i]];
        if (c < 0) return fail(GSM_E_INVALID, "non-ACGT base at offset " + std::to_string(i));
        codes[i] = (uint8_t)c;
        count[c]++;
    }
    return GSM_OK;
}

// BWT of sym[0..n) (sym[n-1] = sentinel 0, bases 1..4) from its suffix array -> packed buckets.
static void bwt_buckets(const uint8_t* sym, const int32_t* SA, uint64_t n, std::vector<uint32_t>& buckets, uint32_t* primary) {
    std::vector<uint8_t> bwt(n);
    for (uint64_t r = 0; r < n; ++r) {
        int32_t p = SA[r];
        if (p == 0) { bwt[r] = 0; *primary = (uint32_t)r; }
        else bwt[r] = (uint8_t)(sym[p - 1] - 1);
    }
    pack_buckets(bwt.data(), n, buckets);
}

static int finish_index(gsm_index* ix, bool have_sa, uint32_t flags) {
    const uint64_t n = ix->n;
    ix->C[0] = 1;
    for (int c = 1; c <= 4; ++c) ix->C[c] = ix->C[c - 1] + ix->count[c - 1];
    std::vector<uint8_t> sym(n);
    for (uint64_t i = 0; i + 1 < n; ++i) sym[i] = ix->codes[i] + 1;
    sym[n - 1] = 0;

    std::thread rev_thread;
    int rev_status = GSM_OK;
    ix->has_rev = (flags & 1u) != 0;
    if (ix->has_rev) {
        rev_thread = std::thread([&]() {
            try {
                std::vector<uint8_t> rsym(n);
                for (uint64_t i = 0; i + 1 < n; ++i) rsym[i] = sym[n - 2 - i];
                rsym[n - 1] = 0;
                std::vector<int32_t> rsa(n);
                suffix_sort(rsym.data(), rsa.data(), (int32_t)n);
                bwt_buckets(rsym.data(), rsa.data(), n, ix->rev, &ix->primary_rev);
            } catch (const std::bad_alloc&) {
                rev_status = GSM_E_NOMEM;
            }
        });
    }
    int status = GSM_OK;
    try {
        if (!have_sa) {
            ix->sa0.resize(n);
            suffix_sort(sym.data(), ix->sa0.data(), (int32_t)n);
        }
        bwt_buckets(sym.data(), ix->sa0.data(), n, ix->fwd, &ix->primary_fwd);
    } catch (const std::bad_alloc&) {
        status = GSM_E_NOMEM;
    }
    if (rev_thread.joinable()) rev_thread.join();
    if (status != GSM_OK || rev_status != GSM_OK) return fail(GSM_E_NOMEM, "out of host memory while building the index");
    return GSM_OK;
}

extern "C" {

int gsm_index_build(const char* bases, uint64_t n_bases, uint32_t flags, gsm_index** out) {
    if (!bases || !out || n_bases == 0) return fail(GSM_E_INVALID, "gsm_index_build: null/empty input");
    if (n_bases + 1 >= (1ull << 31)) return fail(GSM_E_INVALID, "gsm_index_build: n_bases must be < 2^31 - 1 in this build");
    gsm_index* ix = new (std::nothrow) gsm_index();
    if (!ix) return fail(GSM_E_NOMEM, "alloc");
    try {
        ix->n_bases = n_bases;
        ix->n = n_bases + 1;
        int st = decode_bases(bases, n_bases, ix->codes, ix->count);
        if (st == GSM_OK) st = finish_index(ix, false, flags);
        if (st != GSM_OK) { delete ix; return st; }
    } catch (const std::bad_alloc&) {
        delete ix;
        return fail(GSM_E_NOMEM, "out of host memory while building the index");
    }
    *out = ix;
    return GSM_OK;
}

int gsm_index_from_arrays(const char* bases, uint64_t n_bases, const uint32_t* sa1, uint32_t flags, gsm_index** out) {
    if (!bases || !out || !sa1 || n_bases == 0) return fail(GSM_E_INVALID, "gsm_index_from_arrays: null/empty input");
    if (n_bases + 1 >= (1ull << 31)) return fail(GSM_E_INVALID, "n_bases too large");
    gsm_index* ix = new (std::nothrow) gsm_index();
    if (!ix) return fail(GSM_E_NOMEM, "alloc");
    try {
        ix->n_bases = n_bases;
        ix->n = n_bases + 1;
        int st = decode_bases(bases, n_bases, ix->codes, ix->count);
        if (st != GSM_OK) { delete ix; return st; }
        ix->sa0.resize(ix->n);
        std::vector<uint8_t> seen((ix->n + 7) / 8, 0);
        for (uint64_t r = 0; r < ix->n; ++r) {
            uint32_t v = sa1[r];
            if (v < 1 || v > ix->n || (seen[(v - 1) >> 3] >> ((v - 1) & 7)) & 1) {
                delete ix;
                return fail(GSM_E_INVALID, "suffix_array is not a permutation of 1..n");
            }
            seen[(v - 1) >> 3] |= (uint8_t)(1u << ((v - 1) & 7));
            ix->sa0[r] = (int32_t)(v - 1);
        }
        st = finish_index(ix, true, flags);
        if (st != GSM_OK) { delete ix; return st; }
    } catch (const std::bad_alloc&) {
        delete ix;
        return fail(GSM_E_NOMEM, "out of host memory");
    }
    *out = ix;
    return GSM_OK;
}

int gsm_index_info_get(const gsm_index* ix, gsm_index_info* o) {
    if (!ix || !o) return fail(GSM_E_INVALID, "null");
    memset(o, 0, sizeof(*o));
    o->n_bases = ix->n_bases;
    o->n_rows = ix->n;
    o->n_buckets = ix->n / GSM_BUCKET_SYMS + 1;
    o->bucket_bytes = o->n_buckets * GSM_BUCKET_BYTES;
    o->text_words = (ix->n_bases + 15) / 16 + 2;
    for (int c = 0; c < 4; ++c) o->count[c] = ix->count[c];
    for (int c = 0; c < 5; ++c) o->C[c] = ix->C[c];
    o->primary_fwd = ix->primary_fwd;
    o->primary_rev = ix->primary_rev;
    o->has_reverse = ix->has_rev ? 1 : 0;
    return GSM_OK;
}

int gsm_index_export(const gsm_index* ix, uint32_t* sa1, char* bwt) {
    if (!ix) return fail(GSM_E_INVALID, "null");
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t r = 0; r < ix->n; ++r) {
        int32_t p = ix->sa0[r];
        if (sa1) sa1[r] = (uint32_t)p + 1;
        if (bwt) bwt[r] = p == 0 ? '$' : L[ix->codes[p - 1]];
    }
    return GSM_OK;
}

int gsm_index_pack(const gsm_index* ix, void* fwd, void* rev, uint32_t* sa, uint32_t* text2bit) {
    if (!ix) return fail(GSM_E_INVALID, "null");
    if (fwd) memcpy(fwd, ix->fwd.data(), ix->fwd.size() * 4);
    if (rev) {
        if (!ix->has_rev) return fail(GSM_E_INVALID, "index was built without the reverse BWT");
        memcpy(rev, ix->rev.data(), ix->rev.size() * 4);
    }
    if (sa)
        for (uint64_t r = 0; r < ix->n; ++r) sa[r] = (uint32_t)ix->sa0[r] + 1;
    if (text2bit) {
        uint64_t words = (ix->n_bases + 15) / 16;
        text2bit[words] = 0; text2bit[words + 1] = 0;   // two readable pad words (kmer_code)
        for (uint64_t w = 0; w < words; ++w) {
            uint32_t v = 0;
            uint64_t base = w * 16;
            for (uint64_t t = 0; t < 16 && base + t < ix->n_bases; ++t) v |= (uint32_t)ix->codes[base + t] << (30 - 2 * t);
            text2bit[w] = v;
        }
    }
    return GSM_OK;
}

void gsm_index_free(gsm_index* ix) { delete ix; }

int gsm_pack_reads(const char* bases, const uint32_t* lens, uint64_t n_reads, uint32_t* chunk_off, void* packed) {
    if (!bases || !lens || !chunk_off) return fail(GSM_E_INVALID, "null");
    uint64_t off = 0;
    for (uint64_t i = 0; i < n_reads; ++i) {
        chunk_off[i] = (uint32_t)off;
        off += ((uint64_t)lens[i] + 63) / 64;
        if (off >= (1ull << 32)) return fail(GSM_E_CAPACITY, "read batch too large for 32-bit chunk offsets");
    }
    chunk_off[n_reads] = (uint32_t)off;
    if (!packed) return GSM_OK;
    uint32_t* w = (uint32_t*)packed;
    // offsets of every read in `bases`, then the reads are packed by all cores (each read touches only its own chunks)
    std::vector<uint64_t> pos(n_reads + 1, 0);
    for (uint64_t i = 0; i < n_reads; ++i) pos[i + 1] = pos[i] + lens[i];
    unsigned T = std::max(1u, std::thread::hardware_concurrency());
    if (pos[n_reads] < (1u << 20)) T = 1;
    T = (unsigned)std::min<uint64_t>(T, n_reads ? n_reads : 1);
    std::vector<uint64_t> bad(T, UINT64_MAX);
    auto work = [&](unsigned t) {
        const uint64_t a = n_reads * t / T, b = n_reads * (t + 1) / T;
        for (uint64_t i = a; i < b; ++i) {
            uint32_t* dst = w + (uint64_t)chunk_off[i] * 4;
            memset(dst, 0, (size_t)(chunk_off[i + 1] - chunk_off[i]) * 16);
            const char* src = bases + pos[i];
            for (uint32_t k = 0; k < lens[i]; ++k) {
                uint32_t c;
                switch (src[k]) {
                    case 'A': c = 0; break;
                    case 'C': c = 1; break;
                    case 'G': c = 2; break;
                    case 'T': c = 3; break;
                    default: if (bad[t] == UINT64_MAX) bad[t] = i; c = 0; break;
                }
                dst[k >> 4] |= c << (30 - 2 * (k & 15));
            }
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    uint64_t first_bad = UINT64_MAX;
    for (unsigned t = 0; t < T; ++t) first_bad = std::min(first_bad, bad[t]);
    if (first_bad != UINT64_MAX) return fail(GSM_E_INVALID, "non-ACGT base in read " + std::to_string(first_bad));
    return GSM_OK;
}

const char* gsm_last_error(void) { return g_last_error.c_str(); }
int gsm_version(void) { return 100; }

}  // extern "C"
