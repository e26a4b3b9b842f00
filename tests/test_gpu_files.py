"""SURVEY 8f N3 on the GPU: the reference's <ref>-LUT.json through LUT.load_lut / save_lut (LUT.py:50-63), and the single
memory-mappable index file (engine.IndexFile) that replaces the reference's three artefacts: loading it must skip every
rebuild and give identical records."""
import gzip
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from tests import golden_util as gu  # noqa: E402


def _ref_lut_json():
    with gzip.open(os.path.join(gu.GOLDEN, "ref_medium_data-LUT.json.gz"), "rt") as f:
        return f.read()


def test_lut_json_of_the_reference_loads_and_save_lut_writes_the_same_file(tmp_path, monkeypatch):
    import genie_smem_b200 as gs
    g = gu.load_index("medium_data")
    os.makedirs(tmp_path / "data")
    monkeypatch.chdir(tmp_path)
    with open("data/medium_data-LUT.json", "w") as f:
        f.write(_ref_lut_json())
    m = gs.ExactMatch.from_text(g["text"], suffix_array=g["suffix_array"], name="medium_data.fa")
    s = gs.SMEM(m)                                          # SMEM.__init__ loads <ref>-LUT.json eagerly (SMEM.py:11-12)
    assert s.lut.lut_size == 6
    ref = json.loads(_ref_lut_json())
    assert len(s.lut.lut) == len(ref["lut"])
    for key in list(ref["lut"])[::97]:
        assert [list(s.lut.lut[key][0]), s.lut.lut[key][1]] == ref["lut"][key]
    os.rename("data/medium_data-LUT.json", "data/reference-LUT.json")
    s.lut.save_lut()
    assert json.load(open("data/medium_data-LUT.json")) == ref                 # intervals AND positions, every k-mer
    # a LUT file of another reference is refused
    other = json.loads(_ref_lut_json())
    k0 = next(iter(other["lut"]))
    other["lut"][k0][0][0] += 1
    with open("data/medium_data-LUT.json", "w") as f:
        json.dump(other, f)
    with pytest.raises(ValueError):
        gs.LUT(m).load_lut()


def test_index_file_round_trip_skips_every_rebuild(tmp_path):
    import bench
    import genie_smem_b200 as gs
    L = 151
    rng = np.random.default_rng(12)
    ref = rng.integers(0, 4, 1_500_000, dtype=np.uint8)
    starts = rng.integers(0, len(ref) - L, 4000)
    reads = ref[starts[:, None] + np.arange(L)[None, :]].copy()
    mut = rng.random(reads.shape) < 0.012
    reads[mut] = (reads[mut] + 1) & 3
    reads[:300] = rng.integers(0, 4, (300, L), dtype=np.uint8)
    idx = gs.DeviceIndex.build_on_device(ref).build_seed_table()
    idx.build_sampled_sa(32)
    lut = gs.lut_build(idx, 9)
    rmi = bench.train_rmi(idx, 11, (32, 2048), idx.device)      # probe table + None rows
    batch = gs.ReadBatch.from_codes(reads, L)
    e = gs.Engine(idx, len(reads), L, mems_per_read=64, recs_per_read=64)
    want = {}
    for name, method, kw in (("bwa", gs.METHOD_BWA, {"min_len": 1}), ("lut", gs.METHOD_LUT, {"K": 9, "lut": lut}), ("rmi", gs.METHOD_RMI, {"rmi": rmi})):
        r = e.run(method, batch, **kw)
        want[name] = (r.records, r.offsets, r.status)
    path = str(tmp_path / "ref.gsmi")
    size = gs.IndexFile.save(path, idx, lut=lut, lut_K=9, rmi=rmi)
    assert size == os.path.getsize(path)
    head = gs.IndexFile.header(path)
    assert {d["name"] for d in head["sections"]} == {"fwd", "rev", "sa", "text", "ssa", "seed_table", "lut", "rmi_params", "rmi_probe"}
    assert all(d["offset"] % 4096 == 0 for d in head["sections"]) and head["seed_K"] == idx.seed_K and head["lut_K"] == 9
    idx2, lut2, lut_K, rmi2 = gs.IndexFile.load(path, verify=True)
    assert lut_K == 9 and idx2.seed_K == idx.seed_K and rmi2.K == 11 and rmi2.c.n_none_rows == 11 and rmi2.probe is not None
    assert np.array_equal(idx2.suffix_array_host(), idx.suffix_array_host())
    rows = rng.integers(0, idx.n_rows, 1000).astype(np.uint32)
    keep, idx2.sa = idx2.sa, None
    idx2._bind()
    assert np.array_equal(idx2.locate(rows), idx.locate(rows))              # sampled suffix array came along
    idx2.sa = keep
    idx2._bind()
    e2 = gs.Engine(idx2, len(reads), L, mems_per_read=64, recs_per_read=64)
    for name, method, kw in (("bwa", gs.METHOD_BWA, {"min_len": 1}), ("lut", gs.METHOD_LUT, {"K": lut_K, "lut": lut2}), ("rmi", gs.METHOD_RMI, {"rmi": rmi2})):
        r = e2.run(method, batch, **kw)
        assert np.array_equal(r.records, want[name][0]) and np.array_equal(r.offsets, want[name][1]) and np.array_equal(r.status, want[name][2]), name
    # corruption is detected
    sec = next(d for d in head["sections"] if d["name"] == "seed_table")
    with open(path, "r+b") as f:
        f.seek(sec["offset"] + 1234)
        b = f.read(1)
        f.seek(sec["offset"] + 1234)
        f.write(bytes([b[0] ^ 0x40]))
    with pytest.raises(ValueError):
        gs.IndexFile.load(path, verify=True)
