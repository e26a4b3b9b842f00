"""The C-ABI library loads and exports every symbol include/genie_smem.h declares; host-side entry
points work without a GPU; device entry points refuse (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "genie_smem.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gsm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from genie_smem_b200 import _capi as capi
    names = declared_functions()
    assert len(names) >= 18
    lib = C.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/genie_smem.h but not exported"
    assert set(capi.EXPORTS) == set(names), "ctypes bindings out of sync with the header"
    assert capi.lib.gsm_version() >= 100


def test_struct_layouts_match_header():
    from genie_smem_b200 import _capi as capi
    assert C.sizeof(capi.IndexInfo) == 96      # 92 bytes of fields, 8-byte aligned
    assert C.sizeof(capi.DevIndex) == 2 * 8 + 4 * 8 + 5 * 4 + 3 * 4 + 8      # + seed_table pointer
    assert C.sizeof(capi.DevReads) == 8 + 3 * 8 + 2 * 4
    assert C.sizeof(capi.Workspace) == 15 * 8
    assert C.sizeof(capi.DevRmi) == 2 * 4 + 5 * 8 + 2 * 4 + 8 + 8 + 2 * 4      # ... bounds, hazard_slots, hazard_n_slots, reserved0
    from genie_smem_b200.engine import RECORD_DTYPE
    assert RECORD_DTYPE.itemsize == 16


@pytest.mark.parametrize("name", ["small_data", "medium_data", "big_data"])
def test_host_builder_reproduces_reference_arrays(name):
    import genie_smem_b200 as g
    gi = gu.load_index(name)
    h = g.HostIndex.build(gi["text"])
    sa, bwt = h.export()
    assert np.array_equal(sa, gi["suffix_array"])
    assert bwt.decode() == gi["bwt"]
    assert h.count_dic() == gu.meta()[name]["count_dic"]
    # importer ("same index arrays"): identical device layouts from reference-built SA
    h2 = g.HostIndex.from_arrays(gi["text"], gi["suffix_array"])
    for a, b in zip(h.pack(), h2.pack()):
        assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        g.HostIndex.from_arrays(gi["text"], gi["suffix_array"][::-1].copy() * 0 + 1)


def test_host_builder_on_repetitive_and_tiny_texts():
    import genie_smem_b200 as g
    from oracle import ref_port as rp
    rng = np.random.default_rng(0)
    texts = ["A", "AC", "ACGT", "AAAAAAAA", "ACACACACACAC", "TAACCC" * 50, "GATTACA" * 9 + "T"]
    texts += ["".join("ACGT"[c] for c in rng.integers(0, 4, int(n))) for n in rng.integers(1, 400, 20)]
    texts += ["".join("AC"[c] for c in rng.integers(0, 2, 300))]
    for t in texts:
        sa, bwt = g.HostIndex.build(t).export()
        o = rp.RefIndex(t)
        assert np.array_equal(sa.astype(np.int64), o.suffix_array), t
        assert bwt.decode() == o.bwt


def test_packed_index_round_trip(tmp_path):
    """The binary on-disk index (replacement of <ref>-FM.json) reloads to identical device layouts."""
    import genie_smem_b200 as g
    gi = gu.load_index("medium_data")
    h = g.HostIndex.build(gi["text"])
    p = g.PackedIndex.from_host(h)
    p.save(str(tmp_path / "idx"))
    for mmap in (False, True):
        q = g.PackedIndex.load(str(tmp_path / "idx"), mmap=mmap)
        for name in ("fwd", "rev", "sa", "text"):
            assert np.array_equal(getattr(p, name), getattr(q, name))
        assert [int(x) for x in q.info.C] == [int(x) for x in h.info.C]
        assert (q.info.n_rows, q.info.primary_fwd, q.info.primary_rev) == (h.info.n_rows, h.info.primary_fwd, h.info.primary_rev)
    assert np.array_equal(p.sa, gi["suffix_array"])


def test_non_acgt_is_rejected():
    import genie_smem_b200 as g
    with pytest.raises(KeyError):
        g.HostIndex.build("ACGTN")
    with pytest.raises(ValueError):
        g.ReadBatch.from_strings(["ACGT", "acgt"])


def test_read_packing_roundtrip():
    import genie_smem_b200 as g
    reads = ["ACGT", "T" * 65, "", "GATTACA" * 30]
    b = g.ReadBatch.from_strings(reads)
    assert list(b.chunk_off_host) == [0, 1, 3, 3, 7]
    words = b.packed_host.view(np.uint32)
    for i, r in enumerate(reads):
        w = words[b.chunk_off_host[i] * 4:]
        got = "".join("ACGT"[(int(w[p >> 4]) >> (30 - 2 * (p & 15))) & 3] for p in range(len(r)))
        assert got == r


def test_device_entry_points_refuse_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import genie_smem_b200 as g
    from genie_smem_b200 import _capi as capi
    with pytest.raises(g.GsmError) as e:
        g.DeviceIndex(g.HostIndex.build("ACGTACGT"))
    assert e.value.code == capi.E_NODEVICE
    wi = capi.WorkspaceInfo()
    assert capi.lib.gsm_smem_workspace_info(10, 100, C.byref(wi)) == capi.E_NODEVICE
    assert b"no CPU fallback" in capi.lib.gsm_last_error()


def test_hazard_hash_is_an_exact_set():
    """gsm_rmi_hazard_hash (host only): the open-addressing table holds exactly the given codes, stays at most half full, and
    refuses sizes that could not terminate a probe sequence."""
    from genie_smem_b200 import _capi as capi
    from tests.emu.harness import build as build_emu
    emu = C.CDLL(build_emu())
    emu.emu_hz_contains.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    rng = np.random.default_rng(4)
    for n in (0, 1, 7, 511, 5000):
        codes = np.unique(rng.integers(0, 1 << 30, n).astype(np.uint32))
        if n >= 7:
            codes[:3] = (0, 1, (1 << 30) - 1)
            codes = np.unique(codes)
        n_slots = 1024
        while n_slots < 4 * len(codes) + 4:
            n_slots *= 2
        slots = np.zeros(n_slots, np.uint32)
        capi.check(capi.lib.gsm_rmi_hazard_hash(codes.ctypes.data, len(codes), slots.ctypes.data, n_slots))
        assert sorted(slots[slots != 0xFFFFFFFF].tolist()) == codes.tolist()
        inside = set(codes.tolist())
        for c in list(codes[:200]) + rng.integers(0, 1 << 30, 300).tolist():
            assert bool(emu.emu_hz_contains(slots.ctypes.data, n_slots, int(c))) == (int(c) in inside)
    codes = np.arange(600, dtype=np.uint32)
    slots = np.zeros(1024, np.uint32)
    assert capi.lib.gsm_rmi_hazard_hash(codes.ctypes.data, 600, slots.ctypes.data, 1024) == capi.E_INVALID      # more than half full
    assert capi.lib.gsm_rmi_hazard_hash(codes.ctypes.data, 10, slots.ctypes.data, 1000) == capi.E_INVALID       # not a power of two
    bad = np.asarray([5, 0xFFFFFFFF], np.uint32)
    assert capi.lib.gsm_rmi_hazard_hash(bad.ctypes.data, 2, slots.ctypes.data, 1024) == capi.E_INVALID          # the empty marker


def test_header_is_plain_c_and_agrees_with_the_bindings(tmp_path):
    """include/genie_smem.h compiles as C99 (no C++ or torch types in the boundary), links against the library, and its struct
    sizes are the ones the ctypes bindings assume."""
    import subprocess
    from genie_smem_b200 import _capi as capi
    src = tmp_path / "abi.c"
    src.write_text('#include "genie_smem.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %d\\n", sizeof(gsm_dev_index), sizeof(gsm_dev_reads), sizeof(gsm_record),\n'
                   '    sizeof(gsm_dev_rmi), sizeof(gsm_workspace), sizeof(gsm_workspace_info), (int)gsm_version()); return 0; }\n')
    exe = tmp_path / "abi"
    lib_dir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-lgenie_smem", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.check_output([str(exe)], text=True).split()
    sizes = [int(x) for x in out[:6]]
    assert sizes == [C.sizeof(capi.DevIndex), C.sizeof(capi.DevReads), 16, C.sizeof(capi.DevRmi), C.sizeof(capi.Workspace), C.sizeof(capi.WorkspaceInfo)]
    assert int(out[6]) == capi.lib.gsm_version()
