/*
 * smem_oracle.c -- plain-C CPU restatement of GENIE-SMEM's search path.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call this file.
 *
 * Parity status: PINNED -- tests/test_oracle_c.py checks it against the golden fixtures frozen
 * from the live reference (tests/golden/) and against oracle/ref_port.py.
 *
 * It follows the reference literally, quirks and cost included: every prefix of a forward
 * extension is searched from scratch (SMEM/SMEM.py:431-440), backward extension restarts a
 * full search per key (SMEM.py:402-408), the LUT/RMI frame machine is SMEM.py:49-186/235-379,
 * the RMI last-mile search is RMI_LUT.py:95-184.  Substrings of the query are (start, end)
 * pairs instead of Python strings; a dict keyed by such strings is an array indexed by `end`
 * whenever all keys share their start (forward_extension's result), and the final result dict
 * collapses equal strings with memcmp.
 *
 * Index representation: bwt[] codes (0..3, 4 = '$'), occ checkpoints every 64 rows, 1-based sa[]
 * (ExactMatch.py:29-30).  occ_incl(c, i) below is the reference's occurance_matrix[c][i].
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

typedef struct {
    int64_t n;          /* rows = n_bases + 1 */
    uint8_t* bwt;       /* n codes */
    uint8_t* text;      /* n_bases codes */
    const uint32_t* sa; /* n, 1-based, borrowed */
    uint32_t* ckpt;     /* (n/64 + 2) x 4 exclusive counts */
    int64_t C[4];       /* count_dic */
    int64_t cnt[4];
    /* dense LUT (LUT.py:15-35): lo,cnt per code; built on demand */
    int lut_K;
    uint32_t* lut;
} orc_index;

typedef struct {
    int K, n_levels;
    int level_size[8], level_off[8];
    const double* coef;
    const double* icpt;
} orc_rmi;

typedef struct { int64_t lo, hi; } tup;  /* hi < lo : the reference's -1 */
static const tup MISS = {0, -1};
static int is_miss(tup t) { return t.hi < t.lo; }

static int code_of(char ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1; }

/* ------------------------------------------------------------------ index */
/* The arrays of orc_create are filled by all cores (three passes over n rows; the middle one gathers text[sa - 2] at random
 * and takes minutes single-threaded at 10^9 rows).  Rows are cut into one contiguous slice per thread. */
typedef struct {
    orc_index* ix; const char* text; uint64_t n_bases; int pass; int tid, nthreads;
    int64_t cnt[4];          /* pass 0: base counts of the slice; pass 2: BWT symbol counts of the slice's 64-row blocks */
} create_job;

static void* create_worker(void* arg) {
    create_job* j = (create_job*)arg;
    orc_index* ix = j->ix;
    const int64_t n = ix->n;
    if (j->pass == 0) {                 /* text codes + per-slice base counts */
        uint64_t a = j->n_bases * (uint64_t)j->tid / j->nthreads, b = j->n_bases * (uint64_t)(j->tid + 1) / j->nthreads;
        for (uint64_t i = a; i < b; ++i) { int c = code_of(j->text[i]); ix->text[i] = (uint8_t)c; if (c >= 0) j->cnt[c]++; }
    } else if (j->pass == 1) {          /* bwt[r] = text[sa[r] - 2] (ExactMatch.py:59-64), '$' (code 4) where sa[r] == 1 */
        int64_t a = n * j->tid / j->nthreads, b = n * (j->tid + 1) / j->nthreads;
        for (int64_t r = a; r < b; ++r) { uint32_t p = ix->sa[r]; ix->bwt[r] = p == 1 ? 4 : ix->text[p - 2]; }
    } else {                            /* slices of whole 64-row blocks: count symbols, checkpoints filled after the prefix */
        int64_t nb = (n >> 6) + 1, a = nb * j->tid / j->nthreads, b = nb * (j->tid + 1) / j->nthreads;
        if (j->pass == 2) {
            for (int64_t r = a << 6; r < (b << 6) && r < n; ++r) if (ix->bwt[r] < 4) j->cnt[ix->bwt[r]]++;
        } else {                        /* pass 3: j->cnt holds the counts before the slice */
            uint32_t run[4] = {(uint32_t)j->cnt[0], (uint32_t)j->cnt[1], (uint32_t)j->cnt[2], (uint32_t)j->cnt[3]};
            for (int64_t blk = a; blk < b; ++blk) {
                memcpy(ix->ckpt + 4 * blk, run, sizeof(run));
                for (int64_t r = blk << 6; r < ((blk + 1) << 6) && r < n; ++r) if (ix->bwt[r] < 4) run[ix->bwt[r]]++;
            }
        }
    }
    return 0;
}

static void create_pass(create_job* jobs, int nthreads, int pass) {
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) { jobs[t].pass = pass; pthread_create(&th[t], 0, create_worker, &jobs[t]); }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
}

void* orc_create(const char* text, uint64_t n_bases, const uint32_t* sa1) {
    orc_index* ix = (orc_index*)calloc(1, sizeof(orc_index));
    int64_t n = (int64_t)n_bases + 1;
    ix->n = n; ix->sa = sa1;
    ix->bwt = (uint8_t*)malloc(n);
    ix->text = (uint8_t*)malloc(n_bases ? n_bases : 1);
    ix->ckpt = (uint32_t*)malloc(sizeof(uint32_t) * 4 * (n / 64 + 2));
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    int nthreads = n < (1 << 20) ? 1 : (cores < 1 ? 1 : (cores > 256 ? 256 : (int)cores));
    create_job* jobs = (create_job*)calloc((size_t)nthreads, sizeof(create_job));
    for (int t = 0; t < nthreads; ++t) { jobs[t].ix = ix; jobs[t].text = text; jobs[t].n_bases = n_bases; jobs[t].tid = t; jobs[t].nthreads = nthreads; }
    create_pass(jobs, nthreads, 0);
    for (int t = 0; t < nthreads; ++t) for (int c = 0; c < 4; ++c) { ix->cnt[c] += jobs[t].cnt[c]; jobs[t].cnt[c] = 0; }
    ix->C[0] = 1;
    for (int c = 1; c < 4; ++c) ix->C[c] = ix->C[c - 1] + ix->cnt[c - 1];
    create_pass(jobs, nthreads, 1);
    create_pass(jobs, nthreads, 2);
    int64_t run[4] = {0, 0, 0, 0};      /* exclusive prefix of the slices' symbol counts */
    for (int t = 0; t < nthreads; ++t) for (int c = 0; c < 4; ++c) { int64_t v = jobs[t].cnt[c]; jobs[t].cnt[c] = run[c]; run[c] += v; }
    create_pass(jobs, nthreads, 3);
    { uint32_t last[4] = {(uint32_t)run[0], (uint32_t)run[1], (uint32_t)run[2], (uint32_t)run[3]}; memcpy(ix->ckpt + 4 * ((n >> 6) + 1), last, sizeof(last)); }
    free(jobs);
    return ix;
}

void orc_destroy(void* h) {
    orc_index* ix = (orc_index*)h;
    if (!ix) return;
    free(ix->bwt); free(ix->text); free(ix->ckpt); free(ix->lut); free(ix);
}

/* occurance_matrix[c][i] (inclusive), ExactMatch.py:70-90 */
static int64_t occ_incl(const orc_index* ix, int c, int64_t i) {
    int64_t p = i + 1, b = p >> 6;
    int64_t v = ix->ckpt[4 * b + c];
    for (int64_t r = b << 6; r < p; ++r) v += ix->bwt[r] == c;
    return v;
}

/* exact_match_back_prop over q[a:b) (ExactMatch.py:132-151) */
static tup backprop(const orc_index* ix, const uint8_t* q, int a, int b) {
    int64_t start = 1, end = ix->n;
    for (int p = b - 1; p >= a; --p) {
        int c = q[p];
        int64_t cc = ix->C[c];
        if (start - 1 <= 0) start = cc + 1;
        else start = cc + 1 + occ_incl(ix, c, start - 2);
        end = cc + occ_incl(ix, c, end - 1);
        if (start > end) return MISS;
    }
    tup t = {start - 1, end - 1};
    return t;
}

/* exact_match_back_prop_add_one (ExactMatch.py:155-171) */
static tup add_one(const orc_index* ix, int c, tup prev) {
    int64_t start = prev.lo + 1, end = prev.hi + 1, cc = ix->C[c];
    if (start - 1 <= 0) start = cc + 1;
    else start = cc + 1 + occ_incl(ix, c, start - 2);
    end = cc + occ_incl(ix, c, end - 1);
    if (start > end) return MISS;
    tup t = {start - 1, end - 1};
    return t;
}

void orc_backsearch(void* h, const char* reads, const uint32_t* lens, uint64_t n_reads, int64_t* lo, int64_t* hi) {
    const orc_index* ix = (const orc_index*)h;
    uint64_t* off = (uint64_t*)malloc(sizeof(uint64_t) * (n_reads + 1));
    off[0] = 0;
    for (uint64_t i = 0; i < n_reads; ++i) off[i + 1] = off[i] + lens[i];
    for (int64_t i = 0; i < (int64_t)n_reads; ++i) {
        int L = (int)lens[i];
        uint8_t* q = (uint8_t*)malloc(L ? L : 1);
        for (int t = 0; t < L; ++t) q[t] = (uint8_t)code_of(reads[off[i] + t]);
        tup r = backprop(ix, q, 0, L);
        lo[i] = r.lo; hi[i] = r.hi;
        free(q);
    }
    free(off);
}

/* ------------------------------------------------------------------ forward / backward extension */
/* forward_extension(query, start_index, largest = q[c0:start_index), suffix_tuple) (SMEM.py:425-443).
 * All keys are q[c0:j); val[j] holds the tuple of key end j (has[j] says it exists).
 * Returns the end of the longest matched string ("currentSearch[:-1]" on a miss). */
static int fwd_ext(const orc_index* ix, const uint8_t* q, int L, int c0, int start_index, int seeded, tup seed,
                   tup* val, uint8_t* has) {
    memset(has, 0, (size_t)L + 1);
    if (seeded) { val[start_index] = seed; has[start_index] = 1; }
    int cur_end = start_index;
    for (int i = start_index + 1; i <= L; ++i) {
        cur_end = i;
        tup t = backprop(ix, q, c0, i);        /* full search of the whole string, every time */
        if (is_miss(t)) return i - 1;
        val[i] = t; has[i] = 1;
    }
    return cur_end;
}

typedef struct { int i, j; tup t; int end; } best_t;   /* string q[i:j); j == i means "" */

/* backward_extension(query, start_index = c0, forward_matches) (SMEM.py:389-423); keys are q[c0:j)
 * for j in ascending order with has[j]. */
static best_t bwd_ext(const orc_index* ix, const uint8_t* q, int L, int c0, const tup* val, const uint8_t* has) {
    best_t best = {0, 0, MISS, -1};
    int lf = -1;                                   /* longest forward key end */
    for (int j = c0; j <= L; ++j) {
        if (!has[j]) continue;
        if (lf < 0 || (j - c0) > (lf - c0)) lf = j;
        tup t = MISS;
        int first = 1;
        for (int i = c0 - 1; i >= 0; --i) {
            if (first) { t = backprop(ix, q, i, j); first = 0; }     /* SMEM.py:406 */
            else t = add_one(ix, q[i], t);                            /* SMEM.py:408 */
            if (is_miss(t)) break;
            if ((j - i) > (best.j - best.i)) { best.i = i; best.j = j; best.t = t; best.end = j; }
        }
    }
    if (lf >= 0 && (lf - c0) > (best.j - best.i)) { best.i = c0; best.j = lf; best.t = val[lf]; best.end = lf; }
    return best;
}

/* get_SMEM_at_index (SMEM.py:469-484) */
static best_t smem_at(const orc_index* ix, const uint8_t* q, int L, int p, tup* val, uint8_t* has) {
    int fe = fwd_ext(ix, q, L, p, p, 0, MISS, val, has);
    best_t b = bwd_ext(ix, q, L, p, val, has);
    if ((fe - p) > (b.j - b.i)) { best_t r = {p, fe, val[fe], fe}; return r; }
    return b;
}

/* ------------------------------------------------------------------ result dict */
typedef struct { int i, j; int64_t lo, hi; } rec_t;

static void dict_put(const uint8_t* q, rec_t* recs, int* n, int i, int j, tup t) {
    for (int k = 0; k < *n; ++k)
        if (recs[k].j - recs[k].i == j - i && memcmp(q + recs[k].i, q + i, (size_t)(j - i)) == 0) {
            recs[k].lo = t.lo; recs[k].hi = t.hi;       /* same key: value replaced, position kept */
            return;
        }
    recs[*n].i = i; recs[*n].j = j; recs[*n].lo = t.lo; recs[*n].hi = t.hi;
    (*n)++;
}

/* ------------------------------------------------------------------ LUT (LUT.py) */
static uint64_t kcode(const uint8_t* s, int K) {
    uint64_t v = 0;
    for (int t = 0; t < K; ++t) v = (v << 2) | s[t];
    return v;
}

int orc_build_lut(void* h, int K) {
    orc_index* ix = (orc_index*)h;
    if (K < 1 || K > 14) return -1;
    free(ix->lut);
    uint64_t ncodes = 1ull << (2 * K);
    ix->lut = (uint32_t*)calloc(ncodes * 2, sizeof(uint32_t));
    ix->lut_K = K;
    int64_t nb = ix->n - 1;
    /* rows are sorted by suffix, so every k-mer's rows are contiguous: generate_lut's
     * [exact_match_back_prop(kmer), positions] (LUT.py:33-35) is (first row, count) */
    for (int64_t r = 0; r < ix->n; ++r) {
        int64_t p = (int64_t)ix->sa[r] - 1;
        if (p + K > nb) continue;
        uint64_t c = kcode(ix->text + p, K);
        if (ix->lut[2 * c + 1] == 0) ix->lut[2 * c] = (uint32_t)r;
        ix->lut[2 * c + 1]++;
    }
    return 0;
}

/* ------------------------------------------------------------------ RMI (RMI.py, RMI_LUT.py) */
typedef struct { const orc_index* ix; const orc_rmi* m; int raised; } rmi_ctx;

static double rmi_predict(const orc_rmi* m, uint64_t code) {       /* RMI.py:52-69 */
    volatile double x = (double)code;
    int model = 0;
    double p = 0.0;
    for (int lv = 0; lv < m->n_levels; ++lv) {
        int k = m->level_off[lv] + model;
        volatile double prod = x * m->coef[k];                     /* no FMA: sklearn multiplies, then adds */
        p = prod + m->icpt[k];
        int scale = lv + 1 < m->n_levels ? m->level_size[lv + 1] : 1;
        if (!(p >= 1.0)) model = 0;
        else if (p >= (double)scale) model = scale - 1;
        else model = (int)p;
    }
    return p;
}

/* get_ref_seq (RMI_LUT.py:89-92): 1 = string (code in *out), 0 = None */
static int ref_seq(rmi_ctx* c, int64_t ind, uint64_t* out) {
    const orc_index* ix = c->ix;
    *out = 0;
    if (ind < -ix->n || ind >= ix->n) { c->raised = 1; return 0; }   /* IndexError */
    if (ind < 0) ind += ix->n;                                        /* Python negative index */
    int64_t s = ix->sa[ind];
    if (s - 1 + c->m->K > ix->n - 1) return 0;
    *out = kcode(ix->text + (s - 1), c->m->K);
    return 1;
}

static int64_t floordiv2(int64_t v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

/* binary_search (RMI_LUT.py:95-133); depth > 900 stands for RecursionError */
static int64_t bin_search(rmi_ctx* c, uint64_t q, int64_t lower, int64_t upper, int strict, int depth) {
    if (c->raised) return 0;
    if (depth > 900) { c->raised = 1; return 0; }
    if (lower == upper) return lower;
    uint64_t s; int ok;
    if (upper - lower == 1) {
        if (strict) { ok = ref_seq(c, upper, &s); return (ok && s == q) ? upper : lower; }
        ok = ref_seq(c, lower, &s);
        return (ok && s == q) ? lower : upper;
    }
    int64_t mid = floordiv2(lower + upper);
    uint64_t ms; int mok = ref_seq(c, mid, &ms);
    while (!mok && mid > lower) {
        if (c->raised) return 0;
        mid -= 1;
        mok = ref_seq(c, mid, &ms);
        if (mid == lower) {
            if (strict) { ok = ref_seq(c, upper, &s); return (ok && s == q) ? upper : lower; }
            return (mok && ms == q) ? lower : upper;
        }
    }
    if (c->raised) return 0;
    if (!mok) { c->raised = 1; return 0; }           /* None < str: TypeError */
    if (ms < q || (ms == q && strict)) return bin_search(c, q, mid, upper, strict, depth + 1);
    return bin_search(c, q, lower, mid, strict, depth + 1);
}

/* exponential_search (RMI_LUT.py:136-184) */
static tup exp_search(rmi_ctx* c, uint64_t q, int64_t start) {
    const int64_t n_rows = c->ix->n;       /* len(ref_seq) + 1 == len(suffix_array) */
    int have_l = 0, have_u = 0;
    int64_t lower = 0, upper = 0;
    uint64_t cur; int ok = ref_seq(c, start, &cur);
    while (!ok) { if (c->raised) return MISS; start += 1; ok = ref_seq(c, start, &cur); }
    if (cur < q) { lower = start; have_l = 1; }
    else if (cur > q) { upper = start; have_u = 1; }
    int64_t w = 1;
    if (!have_u) {
        while (start + w < n_rows) {
            int64_t ind = start + w;
            w *= 2;
            uint64_t f; int fok = ref_seq(c, ind, &f);
            while (!fok) { if (c->raised) return MISS; ind += 1; fok = ref_seq(c, ind, &f); }
            if (f > q) { upper = ind; have_u = 1; break; }
            if (f < q) { lower = ind; have_l = 1; }
        }
    }
    w = 1;
    if (!have_l) {
        while (start - w >= 0) {
            int64_t ind = start - w;
            w *= 2;
            uint64_t f; int fok = ref_seq(c, ind, &f);
            while (!fok) { if (c->raised) return MISS; ind -= 1; fok = ref_seq(c, ind, &f); }
            if (f < q) { lower = ind; have_l = 1; break; }
            if (f > q) { upper = ind; have_u = 1; }
        }
    }
    if (!have_l) lower = 0;
    if (!have_u) upper = n_rows - 1;
    tup t;
    t.lo = bin_search(c, q, lower, upper, 0, 0);
    t.hi = bin_search(c, q, lower, upper, 1, 0);
    return t;
}

/* get_suffix_rmi (RMI_LUT.py:67-78) */
static tup rmi_lookup(rmi_ctx* c, uint64_t code, double* pred) {
    double p = rmi_predict(c->m, code);
    if (pred) *pred = p;
    if (!(p > -9.0e18 && p < 9.0e18)) { c->raised = 1; return MISS; }
    return exp_search(c, code, (int64_t)p);
}

int orc_rmi_lookup(void* h, int K, int n_levels, const uint32_t* level_sizes, const double* coef, const double* icpt,
                   uint64_t code, double* pred, int64_t* lo, int64_t* hi) {
    orc_rmi m; memset(&m, 0, sizeof(m));
    m.K = K; m.n_levels = n_levels; m.coef = coef; m.icpt = icpt;
    int off = 0;
    for (int l = 0; l < n_levels; ++l) { m.level_size[l] = (int)level_sizes[l]; m.level_off[l] = off; off += (int)level_sizes[l]; }
    rmi_ctx c = {(const orc_index*)h, &m, 0};
    tup t = rmi_lookup(&c, code, pred);
    *lo = t.lo; *hi = t.hi;
    return c.raised ? -1 : 0;
}

/* ------------------------------------------------------------------ the three entry points */
typedef struct {
    const orc_index* ix;
    int method;             /* 0 BWA, 1 LUT, 2 RMI */
    int K;
    rmi_ctx rmi;
} seed_ctx;

/* LUT: `encoded_sub in self.lut.lut` / lut[...][0] (SMEM.py:67-73); RMI: get_suffix_rmi + hi >= lo (SMEM.py:251-253) */
static int seed_lookup(seed_ctx* s, const uint8_t* q, int c, tup* out) {
    uint64_t code = kcode(q + c, s->K);
    if (s->method == 1) {
        uint32_t lo = s->ix->lut[2 * code], n = s->ix->lut[2 * code + 1];
        out->lo = lo; out->hi = (int64_t)lo + n - 1;
        return n != 0;
    }
    *out = rmi_lookup(&s->rmi, code, 0);
    if (s->rmi.raised) return 0;
    return out->hi >= out->lo;
}

/* check_sequential(get_positions(a), get_positions(b)) (SMEM.py:196-202, 262-265; LUT.py:34) */
static int check_sequential(const orc_index* ix, tup a, tup b) {
    for (int64_t x = a.lo; x <= a.hi; ++x) {
        uint32_t px = ix->sa[x < 0 ? x + ix->n : x];
        for (int64_t y = b.lo; y <= b.hi; ++y)
            if (px + 1 == ix->sa[y < 0 ? y + ix->n : y]) return 1;
    }
    return 0;
}

#define UPD(I, J, T, E) do { if (!cand_valid || ((J) - (I)) >= (cand.j - cand.i)) { cand_valid = 1; cand.i = (I); cand.j = (J); cand.t = (T); cand.end = (E); } } while (0)

/* get_smems_lut (SMEM.py:20-192) / get_smems_rmi (SMEM.py:206-384) */
static int smems_seeded(seed_ctx* s, const uint8_t* q, int L, rec_t* recs, tup* val, uint8_t* has) {
    const orc_index* ix = s->ix;
    const int K = s->K;
    int n = 0;
    tup st;
    int fe;
    if (seed_lookup(s, q, 0, &st)) fe = fwd_ext(ix, q, L, 0, K, 1, st, val, has);
    else { if (s->rmi.raised) return -1; fe = fwd_ext(ix, q, L, 0, 0, 0, MISS, val, has); }
    dict_put(q, recs, &n, 0, fe, val[fe]);
    int prev_len = fe, e = fe;
    while (e < L) {
        int fstate = 0, pc = 0, pfw = 0;     /* 0 None, 1 (), 2 frame */
        tup pt = MISS;
        int cand_valid = 0;
        best_t cand = {0, 0, MISS, -1};
        const int prev_start = e - prev_len;
        for (int i = 0; i < K; ++i) {
            if (i >= prev_len) continue;
            int c = e - i;
            if (c + K > L) continue;
            tup ht;
            int hit = seed_lookup(s, q, c, &ht);
            if (s->rmi.raised) return -1;
            if (hit) {
                if (fstate == 0) { fstate = 2; pc = c; pfw = 1; pt = ht; }
                else if (fstate == 1) { fstate = 2; pc = c; pfw = 0; pt = ht; }
                else {
                    if (check_sequential(ix, ht, pt)) {                              /* Case 1 */
                        if (pfw) {
                            fwd_ext(ix, q, L, pc, pc + K, 1, pt, val, has);
                            best_t b = bwd_ext(ix, q, L, pc, val, has);
                            UPD(b.i, b.j, b.t, b.end);
                        } else {
                            if (cand_valid && (pc - prev_start) + K < (cand.j - cand.i)) continue;   /* SMEM.py:94-95 */
                            memset(has, 0, (size_t)L + 1); val[pc + K] = pt; has[pc + K] = 1;
                            best_t b = bwd_ext(ix, q, L, pc, val, has);
                            UPD(b.i, b.j, b.t, b.end);
                        }
                    } else {                                                         /* Case 2 */
                        if (pfw) {
                            int f2 = fwd_ext(ix, q, L, pc, pc + K, 1, pt, val, has);
                            UPD(pc, f2, val[f2], f2);
                        } else UPD(c, c + K, ht, c + K);
                    }
                    pc = c; pfw = 0; pt = ht;
                }
            } else {
                if (fstate == 2) {                                                   /* Case 3 */
                    if (pfw) {
                        int f2 = fwd_ext(ix, q, L, pc, pc + K, 1, pt, val, has);
                        UPD(pc, f2, val[f2], f2);
                    } else UPD(pc, pc + K, pt, pc + K);
                }
                fstate = 1;
            }
        }
        if (fstate == 2) {                                                           /* last frame */
            best_t b;
            if (pfw) { fwd_ext(ix, q, L, pc, pc + K, 1, pt, val, has); b = bwd_ext(ix, q, L, pc, val, has); }
            else { memset(has, 0, (size_t)L + 1); val[pc + K] = pt; has[pc + K] = 1; b = bwd_ext(ix, q, L, pc, val, has); }
            UPD(b.i, b.j, b.t, b.end);
        }
        if (!cand_valid) {
            best_t b = smem_at(ix, q, L, e, val, has);
            dict_put(q, recs, &n, b.i, b.j, b.t);
            e = b.end; prev_len = b.j - b.i;
        } else {
            e = cand.end;
            dict_put(q, recs, &n, cand.i, cand.j, cand.t);
            prev_len = cand.j - cand.i;
        }
    }
    return n;
}

/* get_SMEMS (SMEM.py:456-467) */
static int smems_bwa(const orc_index* ix, const uint8_t* q, int L, int min_len, rec_t* recs, tup* val, uint8_t* has) {
    int n = 0, p = 0;
    while (p < L) {
        best_t b = smem_at(ix, q, L, p, val, has);
        if (b.j - b.i >= min_len) dict_put(q, recs, &n, b.i, b.j, b.t);
        p = b.end;
        if (p < 0) break;     /* base absent from the text: the reference misbehaves; stop */
    }
    return n;
}

/*
 * Batch driver.  reads: concatenated ASCII; out: cap_per_read records of 4 x int64 (i, j, lo, hi)
 * per read, in the reference dict's order; counts[r] = number of entries, -1 = the reference
 * raises on this read, -2 = shorter than K.  Returns 0, or -1 on bad input.  n_threads <= 0 uses
 * every online core (pthreads, reads handed out in blocks of 8 from a shared counter).
 */
typedef struct {
    orc_index* ix; const orc_rmi* m; int method, min_len, K;
    const char* reads; const uint32_t* lens; const uint64_t* off; uint64_t n_reads;
    int64_t* out; uint32_t cap; int32_t* counts;
    volatile int64_t next; volatile int bad;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (;;) {
        int64_t r0 = __sync_fetch_and_add(&j->next, 8);
        if (r0 >= (int64_t)j->n_reads) break;
        int64_t r1 = r0 + 8 < (int64_t)j->n_reads ? r0 + 8 : (int64_t)j->n_reads;
        for (int64_t r = r0; r < r1; ++r) {
            int L = (int)j->lens[r];
            uint8_t* q = (uint8_t*)malloc((size_t)L + 1);
            tup* val = (tup*)malloc(sizeof(tup) * ((size_t)L + 2));
            uint8_t* has = (uint8_t*)malloc((size_t)L + 2);
            rec_t* recs = (rec_t*)malloc(sizeof(rec_t) * ((size_t)L + 2));
            int ok = 1;
            for (int t = 0; t < L; ++t) { int c = code_of(j->reads[j->off[r] + t]); if (c < 0) ok = 0; q[t] = (uint8_t)c; }
            int n = 0, absent = 0;
            /* a base that does not occur in the text: count_dic[char] raises KeyError (ExactMatch.py:140) the first time it is
             * queried, and every base of a read is queried (without this the walk below would not advance past it) */
            for (int t = 0; ok && t < L; ++t) if (j->ix->cnt[q[t]] == 0) absent = 1;
            if (!ok) { j->bad = 1; n = 0; }
            else if (absent) n = -1;
            else if (j->method == 0) n = smems_bwa(j->ix, q, L, j->min_len, recs, val, has);
            else if (L < j->K) n = -2;
            else {
                seed_ctx s; s.ix = j->ix; s.method = j->method; s.K = j->K; s.rmi.ix = j->ix; s.rmi.m = j->m; s.rmi.raised = 0;
                n = smems_seeded(&s, q, L, recs, val, has);
            }
            j->counts[r] = n;
            for (int k = 0; k < n && (uint32_t)k < j->cap; ++k) {
                int64_t* o = j->out + ((uint64_t)r * j->cap + k) * 4;
                o[0] = recs[k].i; o[1] = recs[k].j; o[2] = recs[k].lo; o[3] = recs[k].hi;
            }
            free(q); free(val); free(has); free(recs);
        }
    }
    return 0;
}

int orc_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

int orc_smems(void* h, int method, const char* reads, const uint32_t* lens, uint64_t n_reads, int min_len, int K,
              int n_levels, const uint32_t* level_sizes, const double* coef, const double* icpt, int64_t* out,
              uint32_t cap_per_read, int32_t* counts, int n_threads) {
    orc_index* ix = (orc_index*)h;
    orc_rmi m; memset(&m, 0, sizeof(m));
    if (method == 1) { if (ix->lut_K != K) { if (orc_build_lut(ix, K)) return -1; } }
    if (method == 2) {
        m.K = K; m.n_levels = n_levels; m.coef = coef; m.icpt = icpt;
        int off = 0;
        for (int l = 0; l < n_levels; ++l) { m.level_size[l] = (int)level_sizes[l]; m.level_off[l] = off; off += (int)level_sizes[l]; }
    }
    uint64_t* off = (uint64_t*)malloc(sizeof(uint64_t) * (n_reads + 1));
    off[0] = 0;
    for (uint64_t i = 0; i < n_reads; ++i) off[i + 1] = off[i] + lens[i];
    if (n_threads <= 0) n_threads = orc_max_threads();
    if (n_threads > 256) n_threads = 256;
    job_t job = {ix, &m, method, min_len, K, reads, lens, off, n_reads, out, cap_per_read, counts, 0, 0};
    pthread_t th[256];
    for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], 0, worker, &job);
    worker(&job);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], 0);
    free(off);
    return job.bad ? -1 : 0;
}
