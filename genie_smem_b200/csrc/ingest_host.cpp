// Host side of the read ingest: a multi-threaded FASTQ scanner.  The reference reads one sequence per
// file with a Python line loop (SMEM/ExactMatch.py:104-108); here a 4-line FASTQ held in memory is cut into records by
// memchr over per-thread slices of the buffer, and the sequence lines are copied into one contiguous buffer (the input of
// gsm_pack_reads_device / PipelinedEngine.run_ascii) by the same threads.  No base is interpreted here: characters
// outside ACGT are the packer's business (GSM_E_INVALID, the reference's KeyError).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/genie_smem.h"
#include "host_common.hpp"

using namespace gsm;

namespace {

unsigned pick_threads(uint32_t threads, uint64_t n_bytes) {
    if (threads) return (unsigned)std::min<uint64_t>(threads, n_bytes / 64 + 1);     // explicit: honoured down to 64-byte slices
    const unsigned t = std::max(1u, std::thread::hardware_concurrency());
    return (unsigned)std::min<uint64_t>(t, n_bytes / (1u << 20) + 1);                // auto: at least 1 MB per thread
}

template <typename F>
void parallel_for(unsigned n_threads, F f) {
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; ++t) pool.emplace_back(f, t);
    f(0u);
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int gsm_fastq_scan(const char* buf, uint64_t n_bytes, uint64_t* n_records, uint64_t* seq_off, uint32_t* seq_len, uint64_t cap,
                   uint32_t threads) {
    if (!n_records || (!buf && n_bytes)) return fail(GSM_E_INVALID, "gsm_fastq_scan: null");
    *n_records = 0;
    if (n_bytes == 0) return GSM_OK;
    const unsigned T = pick_threads(threads, n_bytes);
    std::vector<uint64_t> lo(T + 1), first_line(T + 1, 0);
    for (unsigned t = 0; t <= T; ++t) lo[t] = n_bytes * t / T;
    // pass 1: newlines per slice -> index of the first line that STARTS in each slice
    std::vector<uint64_t> nl(T, 0);
    parallel_for(T, [&](unsigned t) {
        uint64_t c = 0;
        const char* p = buf + lo[t];
        const char* e = buf + lo[t + 1];
        while (p < e) {
            const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
            if (!q) break;
            ++c;
            p = q + 1;
        }
        nl[t] = c;
    });
    // first_line[t] = newlines before lo[t] = index of the line that contains byte lo[t]
    uint64_t total_nl = 0;
    for (unsigned t = 0; t < T; ++t) { first_line[t] = total_nl; total_nl += nl[t]; }
    const uint64_t n_lines = total_nl + (buf[n_bytes - 1] == '\n' ? 0 : 1);
    if (n_lines % 4) return fail(GSM_E_INVALID, "gsm_fastq_scan: " + std::to_string(n_lines) + " lines is not a multiple of 4 (truncated FASTQ?)");
    *n_records = n_lines / 4;
    if (!seq_off || !seq_len) return GSM_OK;
    if (cap < *n_records) return fail(GSM_E_CAPACITY, "gsm_fastq_scan: seq_off / seq_len too small");
    // pass 2: every slice walks the lines that start inside it
    std::vector<int> bad(T, 0);
    parallel_for(T, [&](unsigned t) {
        uint64_t pos = lo[t], line = first_line[t];
        if (t > 0 && buf[pos - 1] != '\n') {            // the slice begins inside a line that started earlier: skip to its end
            const char* q = (const char*)memchr(buf + pos, '\n', (size_t)(n_bytes - pos));
            if (!q) return;
            pos = (uint64_t)(q - buf) + 1;
            line += 1;                                  // that newline was counted in this slice (it lies at or after lo[t])
        }
        while (pos < n_bytes && pos < lo[t + 1]) {
            const char* q = (const char*)memchr(buf + pos, '\n', (size_t)(n_bytes - pos));
            uint64_t end = q ? (uint64_t)(q - buf) : n_bytes;
            uint64_t stop = end;
            if (stop > pos && buf[stop - 1] == '\r') --stop;
            const uint64_t k = line & 3u;
            if (k == 0 && buf[pos] != '@') bad[t] = 1;
            if (k == 2 && buf[pos] != '+') bad[t] = 1;
            if (k == 1) {
                seq_off[line >> 2] = pos;
                seq_len[line >> 2] = (uint32_t)(stop - pos);
            }
            pos = end + 1;
            ++line;
        }
    });
    for (unsigned t = 0; t < T; ++t)
        if (bad[t]) return fail(GSM_E_INVALID, "gsm_fastq_scan: not a 4-line FASTQ (records must start with '@' and have a '+' line)");
    return GSM_OK;
}

int gsm_fastq_gather(const char* buf, const uint64_t* seq_off, const uint32_t* seq_len, uint64_t n_records, char* out, uint64_t* base_off,
                     uint32_t threads) {
    if (!base_off || (n_records && (!buf || !seq_off || !seq_len))) return fail(GSM_E_INVALID, "gsm_fastq_gather: null");
    base_off[0] = 0;
    for (uint64_t i = 0; i < n_records; ++i) base_off[i + 1] = base_off[i] + seq_len[i];
    if (!out || n_records == 0) return GSM_OK;
    const unsigned T = pick_threads(threads, base_off[n_records]);
    parallel_for(T, [&](unsigned t) {
        const uint64_t a = n_records * t / T, b = n_records * (t + 1) / T;
        for (uint64_t i = a; i < b; ++i) memcpy(out + base_off[i], buf + seq_off[i], seq_len[i]);
    });
    return GSM_OK;
}

}  // extern "C"
