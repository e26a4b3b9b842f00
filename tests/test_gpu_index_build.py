"""GPU index builder (gsm_index_build_device) against the host SA-IS builder, which is itself pinned to the
reference's own index arrays (test_gpu_parity.py::test_index_builder_matches_reference_arrays, tests/golden).
Bit-exact: suffix array, both bucket arrays, C, primary rows."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import golden_util as gu  # noqa: E402


@pytest.fixture(scope="module")
def gs():
    import genie_smem_b200 as g
    return g


def _texts():
    rng = np.random.default_rng(5)
    rnd = lambda n: "".join("ACGT"[c] for c in rng.integers(0, 4, n))  # noqa: E731
    unit = rnd(700)
    out = {
        "one_base": "G",
        "two_bases": "AA",
        "five": "ACGTA",
        "poly_a_33": "A" * 33,
        "poly_a": "A" * 5000,                              # every suffix ties on every key: log2(n/32) doubling rounds
        "poly_t_then_a": "T" * 100 + "A" * 100,
        "tandem": "ACGT" * 3000,
        "tail_a_run": rnd(3000) + "A" * 70,                # short suffixes tie with genuine A-runs: the virtual ranks
        "tail_equals_head": "C" + "A" * 40 + rnd(500) + "C" + "A" * 40,
        "random_200k": rnd(200_000),
        "repeats": unit * 5 + rnd(1000) + unit[:400] * 3 + "A" * 300 + unit,
        "two_letter": "".join("AC"[c] for c in rng.integers(0, 2, 20000)),
    }
    return out


@pytest.mark.parametrize("name", list(_texts().keys()))
def test_device_build_equals_host_build(gs, name):
    text = _texts()[name]
    host = gs.HostIndex.build(text)
    fwd, rev, sa, _ = host.pack(with_sa=True, with_text=False)
    dev = gs.DeviceIndex.build_on_device(text)
    assert np.array_equal(dev.sa.cpu().numpy().view(np.uint32), sa), "suffix array"
    assert np.array_equal(dev.fwd.cpu().numpy().view(np.uint32), fwd), "forward buckets"
    assert np.array_equal(dev.rev.cpu().numpy().view(np.uint32), rev), "reverse buckets"
    for f in ("n_bases", "n_rows", "n_buckets", "bucket_bytes", "primary_fwd", "primary_rev", "has_reverse"):
        assert int(getattr(dev.info, f)) == int(getattr(host.info, f)), f
    assert list(dev.info.count) == list(host.info.count)
    assert list(dev.info.C) == list(host.info.C)
    _, _, _, text_host = host.pack(with_sa=False, with_text=True)
    assert np.array_equal(dev.text.cpu().numpy().view(np.uint32)[: len(text_host)], text_host), "packed text"


@pytest.mark.parametrize("name", ["small_data", "medium_data", "big_data"])
def test_device_build_equals_reference_arrays(gs, name):
    """The arrays the reference's own create_fm_index produced (frozen in tests/golden)."""
    gidx = gu.load_index(name)
    dev = gs.DeviceIndex.build_on_device(gidx["text"])
    assert np.array_equal(dev.suffix_array_host(), gidx["suffix_array"])
    assert dev.count_dic() == gu.meta()[name]["count_dic"]


def test_device_build_from_codes_and_searches(gs):
    """codes input (no ASCII round trip); the built index drives the search kernels like a host-built one."""
    rng = np.random.default_rng(9)
    codes = rng.integers(0, 4, 300_000, dtype=np.uint8)
    text = "".join("ACGT"[c] for c in codes)
    dev = gs.DeviceIndex.build_on_device(codes)
    ref = gs.DeviceIndex(gs.HostIndex.build(text))
    reads = []
    for _ in range(2000):
        p = int(rng.integers(0, len(text) - 151))
        q = list(text[p:p + 151])
        for k in np.nonzero(rng.random(151) < 0.02)[0]:
            q[k] = "ACGT"[("ACGT".index(q[k]) + 1) % 4]
        reads.append("".join(q))
    out = []
    for idx in (dev, ref):
        e = gs.Engine(idx, len(reads), 151)
        r = e.run(gs.METHOD_BWA, gs.ReadBatch.from_strings(reads), min_len=1)
        out.append((r.records.copy(), r.offsets.copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_device_build_rejects_non_acgt(gs):
    with pytest.raises(ValueError):
        gs.DeviceIndex.build_on_device("ACGTNACGT")
    with pytest.raises(KeyError):
        gs.DeviceIndex.build_on_device("ACGTNACGT")


@pytest.mark.parametrize("sample", [1, 5, 32, 64])
def test_sampled_sa_locate_equals_full_sa(gs, sample):
    """k_locate_sampled (LF walk to every sample-th row) == direct suffix-array reads, full SA dropped afterwards."""
    rng = np.random.default_rng(12)
    for text in ("ACGTA", "A" * 300 + "C" + "A" * 77, "".join("ACGT"[c] for c in rng.integers(0, 4, 150_000))):
        idx = gs.DeviceIndex.build_on_device(text)
        sa = idx.suffix_array_host().copy()
        rows = np.arange(len(sa), dtype=np.uint32) if len(sa) < 2000 else rng.integers(0, len(sa), 30_000).astype(np.uint32)
        assert np.array_equal(idx.locate(rows), sa[rows])
        idx.build_sampled_sa(sample, drop_full=True)
        assert idx.sa is None
        assert np.array_equal(idx.locate(rows), sa[rows])
        assert np.array_equal(idx.locate(np.array([len(sa) + 5], np.uint32)), np.array([0], np.uint32))


def test_device_built_index_round_trips_through_disk(gs, tmp_path):
    """PackedIndex.from_device -> save -> load (mmap) -> DeviceIndex: the same arrays, and it searches."""
    rng = np.random.default_rng(4)
    codes = rng.integers(0, 4, 120_000, dtype=np.uint8)
    dev = gs.DeviceIndex.build_on_device(codes)
    gs.PackedIndex.from_device(dev).save(str(tmp_path / "idx"))
    back = gs.DeviceIndex(gs.PackedIndex.load(str(tmp_path / "idx")))
    for name in ("fwd", "rev", "sa"):
        assert np.array_equal(getattr(back, name).cpu().numpy(), getattr(dev, name).cpu().numpy()), name
    n_text = (len(codes) + 15) // 16
    assert np.array_equal(back.text.cpu().numpy()[:n_text], dev.text.cpu().numpy()[:n_text])
    reads = ["".join("ACGT"[c] for c in codes[p:p + 80]) for p in (0, 5000, 119_900)]
    a = gs.backsearch_batch(dev, gs.ReadBatch.from_strings(reads))
    b = gs.backsearch_batch(back, gs.ReadBatch.from_strings(reads))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.all(a[1] >= 1)


@pytest.mark.parametrize("name", ["medium_data", "big_data"])
def test_exactmatch_surface_on_a_device_built_index(gs, name):
    """ExactMatch.from_text(builder="device"): the reference-shaped class over an index built on the GPU exposes the
    reference's own arrays (suffix_array, bwt_array, count_dic) and answers like the golden sets."""
    gidx = gu.load_index(name)
    m = gs.ExactMatch.from_text(gidx["text"], builder="device", name=name + ".fa")
    assert np.array_equal(m.fm_index["suffix_array"], gidx["suffix_array"])
    assert "".join(m.fm_index["bwt_array"]) == gidx["bwt"]
    assert m.fm_index["count_dic"] == gu.meta()[name]["count_dic"]
    host = gs.ExactMatch.from_text(gidx["text"], builder="host", name=name + ".fa")
    q = gidx["text"][100:180]
    assert m.exact_match_back_prop(q) == host.exact_match_back_prop(q)
    assert m.exact_match(q) == host.exact_match(q)
    assert gs.SMEM(m).get_SMEMS(q[:30] + "T" + q[31:], 1) == gs.SMEM(host).get_SMEMS(q[:30] + "T" + q[31:], 1)
