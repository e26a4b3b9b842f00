"""Host-side FASTA / FASTQ readers (genie_smem_b200/ingest.py): pure numpy, no GPU."""
import numpy as np
import pytest

from genie_smem_b200 import ingest


def test_fastq_reader_ragged_crlf_and_policies(tmp_path):
    reads = ["ACGT", "A", "ACGTNACGT", "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", "acgt", "GGCC"]
    p = tmp_path / "r.fq"
    ingest.write_fastq(str(p), reads)
    bases, off, dropped = ingest.read_fastq(str(p), pin=False)
    assert len(dropped) == 0 and list(np.diff(off)) == [len(r) for r in reads]
    assert bytes(bases).decode() == "".join(reads)
    bases, off, dropped = ingest.read_fastq(str(p), n_policy="drop", pin=False)
    assert list(dropped) == [2, 4]                       # N and lower case: outside ACGT, like the reference's KeyError
    assert bytes(bases).decode() == "ACGT" + "A" + "T" * 37 + "GGCC" and off[-1] == len(bases)
    # CRLF line ends and no trailing newline
    q = tmp_path / "crlf.fq"
    q.write_bytes(b"@a\r\nACG\r\n+\r\nIII\r\n@b\r\nTT\r\n+\r\nII")
    bases, off, _ = ingest.read_fastq(str(q), pin=False)
    assert bytes(bases) == b"ACGTT" and list(off) == [0, 3, 5]
    with pytest.raises(ValueError):
        (tmp_path / "bad.fq").write_text("@a\nACG\n+\n")
        ingest.read_fastq(str(tmp_path / "bad.fq"), pin=False)
    (tmp_path / "empty.fq").write_text("")
    bases, off, _ = ingest.read_fastq(str(tmp_path / "empty.fq"), pin=False)
    assert len(bases) == 0 and list(off) == [0]


def test_fasta_reader_matches_reference_parser(tmp_path):
    """Same text as ExactMatch.load_ref_sequence (reference ExactMatch.py:43-50: skip the header, join stripped lines)."""
    seq = "ACGTTGCA" * 40 + "AC"
    p = tmp_path / "ref.fa"
    p.write_text(">chr test\n" + "\n".join(seq[i:i + 60] for i in range(0, len(seq), 60)) + "\n")
    names, seqs = ingest.read_fasta(str(p))
    assert names == ["chr test"] and bytes(seqs[0]).decode() == seq
    with open(p) as f:                      # the reference's own parsing rule
        f.readline()
        assert "".join(line.strip() for line in f) == seq
    m = tmp_path / "multi.fa"
    m.write_text(">a\nAC\nGT\n\n>b\nTTT\n")
    names, seqs = ingest.read_fasta(str(m))
    assert names == ["a", "b"] and [bytes(s).decode() for s in seqs] == ["ACGT", "TTT"]


def test_fastq_scanner_slice_boundaries(tmp_path):
    """The native scanner cuts the buffer into per-thread slices: every slice count must give the same records, whatever
    falls on a boundary (header, sequence, '+', quality, CRLF, a line longer than a slice)."""
    rng = np.random.default_rng(7)
    reads = ["".join("ACGT"[c] for c in rng.integers(0, 4, int(L))) for L in rng.integers(0, 90, 300)]
    reads[17] = "ACGT" * 200                               # longer than a 64-byte slice many times over
    p = tmp_path / "r.fq"
    ingest.write_fastq(str(p), reads, names=["read%d/%s" % (i, "x" * (i % 23)) for i in range(len(reads))])
    crlf = tmp_path / "c.fq"
    crlf.write_bytes(p.read_bytes().replace(b"\n", b"\r\n")[:-2])      # CRLF, no trailing newline
    for path in (p, crlf):
        ref = None
        for threads in (1, 2, 3, 5, 8, 13, 64, 1000):
            bases, off, _ = ingest.read_fastq(str(path), pin=False, threads=threads)
            got = [bytes(bases[off[i]:off[i + 1]]).decode() for i in range(len(off) - 1)]
            assert got == reads, (str(path), threads)
    bad = tmp_path / "b.fq"
    bad.write_text("@a\nACGT\n+\nIIII\nXb\nAC\n+\nII\n")
    with pytest.raises(ValueError):
        ingest.read_fastq(str(bad), pin=False, threads=2)


def test_fastq_cuts_fall_on_record_boundaries():
    """ingest.fastq_cuts (the only host-side work of PipelinedEngine.run_fastq): every cut is the first byte of a record,
    also when quality lines start with '@'."""
    import numpy as np
    from genie_smem_b200 import ingest
    rng = np.random.default_rng(1)
    recs = []
    for i in range(2000):
        L = int(rng.integers(30, 160))
        s = "".join("ACGT"[c] for c in rng.integers(0, 4, L))
        q = "".join(chr(int(c)) for c in rng.integers(33, 75, L))
        if i % 7 == 0:
            q = "@" + q[1:]
        recs.append(f"@r{i} desc\n{s}\n+\n{q}\n")
    buf = np.frombuffer("".join(recs).encode(), np.uint8)
    starts = set(np.cumsum([0] + [len(r) for r in recs]).tolist())
    for k in (1, 2, 5, 16, 100, 5000):
        cuts = ingest.fastq_cuts(buf, k)
        assert cuts[0] == 0 and cuts[-1] == len(buf) and cuts == sorted(set(cuts))
        assert all(c in starts for c in cuts), k
    assert ingest.fastq_cuts(buf[:0], 4) == [0, 0] or ingest.fastq_cuts(buf[:0], 4) == [0]
