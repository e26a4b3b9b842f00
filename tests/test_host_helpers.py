"""Host-side helpers that need no GPU: chunk schedule of the pipelined engine, bench.py argument / workload plumbing,
read-sharding arithmetic."""
import sys
import types

import numpy as np
import pytest


def test_pipelined_chunk_bounds_cover_the_batch():
    from genie_smem_b200.engine import PipelinedEngine
    for c in (1, 7, 64, 1000, 3_125_000):
        o = types.SimpleNamespace(chunk_reads=c)
        for n in (0, 1, c - 1, c, c + 1, 4 * c, 4 * c + 1, 16 * c, 16 * c + 3, 50_000_000):
            if n < 0:
                continue
            b = PipelinedEngine._chunk_bounds(o, n)
            assert b[0] == 0 and b[-1] == n
            assert all(0 < y - x <= c for x, y in zip(b, b[1:])), (c, n)
            if n > 4 * c and c >= 1_000_000:     # ramp: the first and last chunks are an eighth of a full one
                assert b[1] - b[0] == c // 8 and b[-1] - b[-2] == c // 8


def test_bench_configs_and_workload(monkeypatch):
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse_args()
    assert (a.ref_bases, a.total_reads, a.seed, a.scaling) == (1_000_000_000, 50_000_000, 1000, "strong") and "configs[3]" in a.cfg_name
    w = bench.workload_dict(a, 8, 6_250_000)
    assert "x8" in w["parallelism"] and w["read_len"] == 151 and "configs[3]" in w["workload"]
    assert w["reads_per_step"] == 50_000_000 and w["reads_per_gpu"] == 6_250_000          # strong scaling: the config's reads are sharded
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "c3", "--gpus", "2", "--scaling", "weak"])
    a = bench.parse_args()
    assert (a.ref_bases, a.total_reads, a.seed, a.gpus) == (100_000_000, 10_000_000, 100, 2)
    assert bench.workload_dict(a, 2, a.total_reads)["reads_per_step"] == 20_000_000
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "c5"])
    a = bench.parse_args()
    assert (a.ref_bases, a.total_reads) == (3_000_000_000, 100_000_000) and a.experts[1] > bench.CONFIGS["c4"]["experts"][1]
    monkeypatch.setattr(sys, "argv", ["bench.py", "--ref-bases", "2000000", "--reads", "1000"])
    a = bench.parse_args()
    assert a.cfg_name == "custom size" and a.experts[1] >= 1024 and a.total_reads == 1000


def test_bench_parity_comparison_applies_the_reference_dict_semantics():
    """compare_with_oracle: duplicate SMEM strings collapse (first insertion keeps its place, last value wins), raising
    reads must be flagged, any difference counts."""
    import bench
    from genie_smem_b200.engine import RECORD_DTYPE
    reads = np.array([[0, 1, 0, 1, 2], [3, 3, 3, 3, 3]], np.uint8)
    recs = np.array([(0, 0, 2, 5, 6), (0, 2, 4, 7, 8), (0, 4, 5, 1, 1)], RECORD_DTYPE)      # read 0: "AC" twice (second value wins), then "G"
    offs = np.array([0, 3, 3])
    status = np.array([0, 1], np.uint8)
    out = np.zeros((2, 6, 4), np.int64)
    out[0, 0] = (0, 2, 7, 8)
    out[0, 1] = (4, 5, 1, 1)
    counts = np.array([2, -1], np.int32)
    assert bench.compare_with_oracle(reads, (recs, offs, status), out, counts, 2) == (0, 1)
    out[0, 1] = (4, 5, 1, 2)
    assert bench.compare_with_oracle(reads, (recs, offs, status), out, counts, 2) == (1, 1)
    status[1] = 0                                                                            # the kernel missed a raising read
    assert bench.compare_with_oracle(reads, (recs, offs, status), out, counts, 2) == (2, 1)


def test_synthetic_reads_are_substitution_only():
    import bench
    ref = bench.make_reference(50_000, 7)
    r = bench.make_reads_host(ref, 200, 151, seed=8, sub_rate=0.02)
    assert r.shape == (200, 151) and r.dtype == np.uint8 and r.max() <= 3
    again = bench.make_reads_host(ref, 200, 151, seed=8, sub_rate=0.02)
    assert np.array_equal(r, again)                        # the CPU arms regenerate exactly these reads
    exact = bench.make_reads_host(ref, 50, 151, seed=9, sub_rate=0.0)
    text = ref.tobytes()
    for row in exact:
        assert row.tobytes() in text


def test_shard_range_partitions_reads():
    from genie_smem_b200.sharding import shard_range
    for n in (0, 1, 7, 50_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_host_read_packer_multithreaded_equals_bitwise_definition():
    """gsm_pack_reads (all cores above 1 MB of bases): MSB-first 2-bit words, 16-byte aligned reads, zeroed padding; the first
    read with a character outside ACGT is reported."""
    import ctypes as C
    from genie_smem_b200 import _capi as capi
    rng = np.random.default_rng(2)
    lens = rng.integers(0, 400, 9000).astype(np.uint32)
    lens[:5] = [0, 1, 64, 65, 63]
    codes = [rng.integers(0, 4, int(L), dtype=np.uint8) for L in lens]
    joined = b"".join(np.frombuffer(b"ACGT", np.uint8)[c].tobytes() for c in codes)
    assert len(joined) > 1 << 20                                   # the multi-threaded path
    off = np.zeros(len(lens) + 1, np.uint32)
    capi.check(capi.lib.gsm_pack_reads(joined, lens.ctypes.data, len(lens), off.ctypes.data, None))
    assert np.array_equal(np.diff(off.astype(np.int64)), (lens.astype(np.int64) + 63) // 64)
    packed = np.full(int(off[-1]) * 4 + 4, 0xFFFFFFFF, np.uint32)  # dirty buffer: padding must be zeroed by the packer
    capi.check(capi.lib.gsm_pack_reads(joined, lens.ctypes.data, len(lens), off.ctypes.data, packed.ctypes.data))
    for i in (0, 1, 2, 3, 4, 17, 4000, 8999):
        words = packed[int(off[i]) * 4:int(off[i + 1]) * 4]
        exp = np.zeros(len(words), np.uint32)
        for k, c in enumerate(codes[i]):
            exp[k >> 4] |= np.uint32(int(c) << (30 - 2 * (k & 15)))
        assert np.array_equal(words, exp), i
    bad = bytearray(joined)
    start = int(lens[:7000].sum())
    bad[start + 3] = ord("N")
    with pytest.raises(ValueError, match="read 7000"):
        capi.check(capi.lib.gsm_pack_reads(bytes(bad), lens.ctypes.data, len(lens), off.ctypes.data, packed.ctypes.data))


def test_bench_names_the_limiter_from_its_own_numbers():
    """bench.py's `limiter` object on the committed 1- and 8-GPU records: the sweep kernel bounds the device step; end to end the
    host side shows up once the per-rank batch is small.  Never raises on a malformed record."""
    import json
    import os
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, world in (("r02_bench_c4_1gpu.json", 1), ("r02_bench_c4_8gpu.json", 8)):
        d = json.load(open(os.path.join(root, "profiles", name)))
        lim = bench.name_limiter(d["ms_per_step"], d["roofline"]["ms_sweep"], d["roofline"]["ms_select_scan_write"], d["e2e"], 50_000_000, world)
        assert lim["device"]["limiter"].startswith("k_sweep1") and 0.7 < lim["device"]["share_of_step"] <= 1.0
        assert abs(lim["device"]["ms_sweep"] + lim["device"]["ms_select_scan_write_local"] + lim["device"]["ms_rest"] - d["ms_per_step"]) < 0.01
        assert lim["e2e"]["over_device_step"] >= 1.0
    assert "unavailable" in bench.name_limiter(1.0, 0.5, 0.2, None, 10, 1)
