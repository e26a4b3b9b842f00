"""Device-side read ingest (gsm_pack_reads_device, PipelinedEngine.run_ascii) against the host packer and the
packed-input paths: identical packed bytes, identical records."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gs():
    import genie_smem_b200 as g
    return g


@pytest.fixture(scope="module")
def setup(gs):
    rng = np.random.default_rng(21)
    text = "".join("ACGT"[c] for c in rng.integers(0, 4, 400_000))
    idx = gs.DeviceIndex(gs.HostIndex.build(text))
    reads = []
    for _ in range(5000):
        L = 151
        p = int(rng.integers(0, len(text) - L))
        q = list(text[p:p + L])
        for k in np.nonzero(rng.random(L) < 0.015)[0]:
            q[k] = "ACGT"[("ACGT".index(q[k]) + 1) % 4]
        reads.append("".join(q))
    return text, idx, reads


def test_device_pack_fixed_length_equals_host_pack(gs, setup):
    import torch
    _, _, reads = setup
    host = gs.ReadBatch.from_strings(reads)
    codes = np.frombuffer("".join(reads).encode(), np.uint8).reshape(len(reads), -1)
    asc = torch.from_numpy(codes.copy()).cuda()
    dev = gs.ReadBatch.from_device_bases(asc, ascii=True).to_host(pin=False)
    assert np.array_equal(dev.chunk_off_host, host.chunk_off_host)
    assert np.array_equal(dev.len_host, host.len_host)
    assert np.array_equal(dev.packed_host, host.packed_host)
    lut = np.zeros(256, np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    dev2 = gs.ReadBatch.from_device_bases(torch.from_numpy(lut[codes]).cuda()).to_host(pin=False)
    assert np.array_equal(dev2.packed_host, host.packed_host)


def test_device_pack_ragged_equals_host_pack(gs):
    import torch
    rng = np.random.default_rng(3)
    reads = ["".join("ACGT"[c] for c in rng.integers(0, 4, int(L))) for L in rng.integers(1, 400, 3000)]
    reads += ["A", "ACGT" * 16, "T" * 64, "G" * 65, "C" * 63]
    host = gs.ReadBatch.from_strings(reads)
    flat = torch.from_numpy(np.frombuffer("".join(reads).encode(), np.uint8).copy()).cuda()
    off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    dev = gs.ReadBatch.from_device_bases(flat, base_off=off, ascii=True).to_host(pin=False)
    assert np.array_equal(dev.chunk_off_host, host.chunk_off_host)
    assert np.array_equal(dev.len_host, host.len_host)
    assert np.array_equal(dev.packed_host, host.packed_host)


def test_device_pack_rejects_non_acgt(gs):
    import torch
    reads = ["ACGTACGT", "ACGNACGT", "TTTTTTTT"]
    asc = torch.from_numpy(np.frombuffer("".join(reads).encode(), np.uint8).reshape(3, 8).copy()).cuda()
    with pytest.raises(KeyError):
        gs.ReadBatch.from_device_bases(asc, ascii=True)


@pytest.mark.parametrize("method", ["bwa", "lut"])
def test_run_ascii_equals_packed_paths(gs, setup, method):
    import torch
    _, idx, reads = setup
    m = {"bwa": gs.METHOD_BWA, "lut": gs.METHOD_LUT}[method]
    kw = {"min_len": 1} if method == "bwa" else {"K": 8, "lut": gs.lut_build(idx, 8)}
    plain = gs.Engine(idx, len(reads), 151).run(m, gs.ReadBatch.from_strings(reads), **kw)
    pipe = gs.PipelinedEngine(idx, len(reads), 151, n_chunks=5)
    asc = torch.from_numpy(np.frombuffer("".join(reads).encode(), np.uint8).reshape(len(reads), -1).copy()).pin_memory()
    for _ in range(2):                 # second call reuses every cached buffer
        got = pipe.run_ascii(m, asc, 151, **kw)
        assert np.array_equal(got.records, plain.records)
        assert np.array_equal(got.offsets, plain.offsets)
    packed = pipe.run(m, gs.ReadBatch.from_strings(reads, pin=True), **kw)
    assert np.array_equal(packed.records, plain.records)


def test_run_ascii_rejects_non_acgt(gs, setup):
    import torch
    _, idx, reads = setup
    bad = list(reads[:600])
    bad[417] = bad[417][:70] + "N" + bad[417][71:]
    asc = torch.from_numpy(np.frombuffer("".join(bad).encode(), np.uint8).reshape(len(bad), -1).copy()).pin_memory()
    pipe = gs.PipelinedEngine(idx, len(bad), 151, n_chunks=4)
    with pytest.raises(KeyError, match="417"):
        pipe.run_ascii(gs.METHOD_BWA, asc, 151)


def test_fastq_file_to_records(gs, setup, tmp_path):
    """FASTQ on disk -> host reader -> GPU packing -> SMEM records, ragged reads and the 'drop' policy included."""
    import torch
    text, idx, reads = setup
    rng = np.random.default_rng(6)
    ragged = [r[: int(rng.integers(20, 152))] for r in reads[:700]]
    ragged[13] = ragged[13][:5] + "N" + ragged[13][6:]
    path = str(tmp_path / "reads.fq")
    gs.write_fastq(path, ragged)
    bases, off, dropped = gs.read_fastq(path, n_policy="drop")
    assert list(dropped) == [13]
    kept = [r for i, r in enumerate(ragged) if i != 13]
    batch = gs.ReadBatch.from_device_bases(torch.from_numpy(bases).cuda(), base_off=off, ascii=True)
    e = gs.Engine(idx, len(kept), 151)
    batch.to_host(pin=False)
    got = e.run(gs.METHOD_BWA, batch, min_len=1)
    exp = gs.Engine(idx, len(kept), 151).run(gs.METHOD_BWA, gs.ReadBatch.from_strings(kept), min_len=1)
    assert np.array_equal(got.records, exp.records) and np.array_equal(got.offsets, exp.offsets)
    bases, off, _ = gs.read_fastq(path, n_policy="error")
    with pytest.raises(KeyError):
        gs.ReadBatch.from_device_bases(torch.from_numpy(bases).cuda(), base_off=off, ascii=True)


def _fastq_bytes(reads, rng, crlf=False, final_newline=True):
    nl = "\r\n" if crlf else "\n"
    recs = []
    for i, r in enumerate(reads):
        q = "".join(chr(int(c)) for c in rng.integers(33, 74, len(r)))
        if i % 5 == 0 and len(q):
            q = "@" + q[1:]                       # quality lines may start with '@': the record cutter must not be fooled
        recs.append(f"@read{i} some description{nl}{r}{nl}+{nl}{q}{nl}")
    data = "".join(recs)
    if not final_newline:
        data = data[:-len(nl)]
    return data.encode()


@pytest.mark.parametrize("crlf,final_newline,n_chunks", [(False, True, None), (True, True, 3), (False, False, 7)])
def test_fastq_bytes_cut_and_packed_on_the_gpu_equal_the_string_path(gs, setup, crlf, final_newline, n_chunks):
    """PipelinedEngine.run_fastq (file bytes in: records cut and 2-bit packed by GPU kernels) == Engine.run on the same reads
    as Python strings.  Ragged read lengths, CRLF line ends, a missing final line feed, '@' at the start of quality lines."""
    import torch
    text, idx, reads = setup
    rng = np.random.default_rng(4)
    mixed = []
    for k, r in enumerate(reads[:3000]):
        mixed.append(r[: int(rng.integers(20, 152))] if k % 3 else r)
    data = _fastq_bytes(mixed, rng, crlf, final_newline)
    fq = torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory()
    plain = gs.Engine(idx, len(mixed), 160, mems_per_read=48, recs_per_read=48)
    want = plain.run(gs.METHOD_BWA, gs.ReadBatch.from_strings(mixed), min_len=1)
    pipe = gs.PipelinedEngine(idx, len(mixed) + 10, 160, n_chunks=2, mems_per_read=48, recs_per_read=48)
    got = pipe.run_fastq(gs.METHOD_BWA, fq, min_len=1, n_chunks=n_chunks)
    assert len(got.offsets) == len(mixed) + 1
    assert np.array_equal(got.offsets, want.offsets) and np.array_equal(got.records, want.records) and np.array_equal(got.status, want.status)
    lut = gs.lut_build(idx, 8)
    got = pipe.run_fastq(gs.METHOD_LUT, fq, K=8, lut=lut)
    want = plain.run(gs.METHOD_LUT, gs.ReadBatch.from_strings(mixed), K=8, lut=lut)
    assert np.array_equal(got.offsets, want.offsets) and np.array_equal(got.records, want.records) and np.array_equal(got.status, want.status)


def test_fastq_on_the_gpu_reports_bad_input(gs, setup):
    import torch
    text, idx, reads = setup
    rng = np.random.default_rng(5)
    pipe = gs.PipelinedEngine(idx, 1000, 160, n_chunks=2, mems_per_read=48, recs_per_read=48)

    def run(data):
        return pipe.run_fastq(gs.METHOD_BWA, torch.frombuffer(bytearray(data), dtype=torch.uint8).pin_memory(), min_len=1)

    good = _fastq_bytes(reads[:200], rng)
    assert len(run(good).offsets) == 201
    bad_base = _fastq_bytes(reads[:100] + [reads[100][:70] + "N" + reads[100][71:]] + reads[101:200], rng)
    with pytest.raises(KeyError):                              # BaseError is a KeyError (the reference's, ExactMatch.py:139) and a ValueError
        run(bad_base)
    lines = good.decode().split("\n")
    lines[4 * 50 + 2] = "-"                                     # third line of record 50 does not start with '+'
    with pytest.raises(ValueError):
        run("\n".join(lines).encode())
    with pytest.raises(ValueError):
        run(good[: len(good) // 2 + 7])                         # truncated in the middle of a record
    with pytest.raises(Exception):
        gs.PipelinedEngine(idx, 50, 160, n_chunks=2).run_fastq(gs.METHOD_BWA, torch.frombuffer(bytearray(good), dtype=torch.uint8).pin_memory())
