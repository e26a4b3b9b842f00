"""TEST INFRASTRUCTURE: host build of the kernels' per-read logic (see emu_logic.cpp).

Lets the GPU-less container check sweep_logic.cuh / select_logic.cuh / fm_core.cuh -- the exact
headers the CUDA kernels are compiled from -- against the oracle.  Never imported by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libemu_logic.so")
SRC = os.path.join(HERE, "emu_logic.cpp")
CSRC = os.path.join(HERE, "..", "..", "genie_smem_b200", "csrc")


def build():
    os.makedirs(BUILD, exist_ok=True)
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("fm_core.cuh", "sweep_logic.cuh", "select_logic.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", SRC, "-o", LIB])
    return LIB


class EmuIndex(C.Structure):
    _fields_ = [("fwd", C.c_void_p), ("rev", C.c_void_p), ("sa", C.c_void_p), ("text", C.c_void_p), ("C", C.c_uint32 * 4),
                ("cnt", C.c_uint32 * 4), ("prim_f", C.c_uint32), ("prim_r", C.c_uint32), ("n_rows", C.c_uint32), ("uniq_mode", C.c_uint32),
                ("n_bases", C.c_uint64), ("isa", C.c_void_p)]


class Emu:
    """Host arrays of one index (built through the product's own host builder) + emu entry points."""

    def __init__(self, text, sa_1based=None):
        from genie_smem_b200 import _capi as capi
        self.lib = C.CDLL(build())
        P, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        self.lib.emu_sweep.argtypes = [C.POINTER(EmuIndex), P, u32, P, u32, C.POINTER(u64), u32, P]
        self.lib.emu_smem.argtypes = [C.POINTER(EmuIndex), C.c_int, P, u32, u32, u32, P, u32, P, P, P, P, u32, u32, P, u32, P]
        self.lib.emu_rmi_fast.argtypes = [C.POINTER(EmuIndex), u32, u32, P, P, P, u32, P, u64, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                          C.POINTER(u32), C.POINTER(u32)]
        self.lib.emu_rmi_arith.argtypes = [C.POINTER(EmuIndex), u32, u32, P, P, P, u32, P, u64, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                           C.POINTER(u32)]
        self.lib.emu_bwa_picks.argtypes = [u32, P, P]
        self.lib.emu_bwa_picks.restype = u32
        self.lib.emu_rmi_arith_fuzz.argtypes = [u32, u32, P, u64, u64]
        self.lib.emu_rmi_arith_fuzz.restype = u64
        self.lib.emu_rmi_search.argtypes = [C.POINTER(EmuIndex), u32, u32, P, P, P, u64, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                            C.POINTER(u32)]
        self.lib.emu_rmi_hazards.argtypes = [C.POINTER(EmuIndex), u32, u32, P, P, P, u32, P, P, u64]
        self.lib.emu_rmi_hazards.restype = u64
        self.lib.emu_hz_build.argtypes = [P, u64, P, u32]
        self.lib.emu_hz_contains.argtypes = [P, u32, u32]
        self.lib.emu_read_hazard_free.argtypes = [P, u32, u32, P, u32]
        self.lib.emu_kmer_code.argtypes = [P, u32, u32]
        self.lib.emu_kmer_code.restype = u64
        self.lib.emu_locate.argtypes = [C.POINTER(EmuIndex), u32, u64, P, P]
        self.lib.emu_seed_build.argtypes = [C.POINTER(EmuIndex), u32, P]
        self.lib.emu_counters.argtypes = [P, C.c_int]
        self.lib.emu_lut_build.argtypes = [C.POINTER(EmuIndex), u32, P]
        self.lib.emu_rmi_lookup.argtypes = [C.POINTER(EmuIndex), u32, u32, P, P, P, u64, C.POINTER(C.c_double),
                                            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        h = C.c_void_p()
        if sa_1based is None:
            capi.check(capi.lib.gsm_index_build(text.encode(), len(text), 1, C.byref(h)))
        else:
            sa_in = np.ascontiguousarray(sa_1based, dtype=np.uint32)
            capi.check(capi.lib.gsm_index_from_arrays(text.encode(), len(text), sa_in.ctypes.data, 1, C.byref(h)))
        info = capi.IndexInfo()
        capi.check(capi.lib.gsm_index_info_get(h, C.byref(info)))
        self.info = info
        self.fwd = np.zeros(info.n_buckets * 16, np.uint32)
        self.rev = np.zeros(info.n_buckets * 16, np.uint32)
        self.sa = np.zeros(info.n_rows, np.uint32)
        self.text = np.zeros(info.text_words, np.uint32)
        capi.check(capi.lib.gsm_index_pack(h, self.fwd.ctypes.data, self.rev.ctypes.data, self.sa.ctypes.data, self.text.ctypes.data))
        capi.lib.gsm_index_free(h)
        self.capi = capi
        e = EmuIndex()
        e.fwd, e.rev, e.sa, e.text = self.fwd.ctypes.data, self.rev.ctypes.data, self.sa.ctypes.data, self.text.ctypes.data
        for c in range(4):
            e.C[c] = info.C[c]
            e.cnt[c] = info.count[c]
        e.prim_f, e.prim_r, e.n_rows, e.n_bases = info.primary_fwd, info.primary_rev, info.n_rows, info.n_bases
        self.e = e
        self.n_rows = int(info.n_rows)
        self._lut = {}
        self._seed = {}
        self.seed_K = 0          # seed table used by sweep()/smem(); 0 = plain stepping
        self.isa = np.zeros(self.n_rows, np.uint32)                      # inverse suffix array: row of the suffix at each text index
        self.isa[self.sa.astype(np.int64) - 1] = np.arange(self.n_rows, dtype=np.uint32)
        self.e.isa = None
        self.rmi_fast = False    # smem(2, ...): try the error-bounded search (rmi_fast_lookup) before the literal one

    @property
    def uniq(self):
        """Unique-match shortcut of the sweep logic: False / 0 off, 1 = forward only (what k_sweep1 runs: suffix array +
        text), True / 2 = forward and backward (also needs the inverse suffix array)."""
        return int(self.e.uniq_mode)

    @uniq.setter
    def uniq(self, on):
        mode = 2 if on is True else int(on)
        self.e.uniq_mode = mode
        self.e.isa = self.isa.ctypes.data if mode == 2 else None

    def seed_table(self, K):
        if K not in self._seed:
            t = np.zeros(4 << (2 * K), np.uint32)
            self.lib.emu_seed_build(C.byref(self.e), K, t.ctypes.data)
            self._seed[K] = t
        return self._seed[K]

    def _seed_args(self):
        if not self.seed_K:
            return 0, None
        return self.seed_K, self.seed_table(self.seed_K).ctypes.data

    def counters(self, reset=True):
        out = np.zeros(8, np.uint64)
        self.lib.emu_counters(out.ctypes.data, 1 if reset else 0)
        return out

    def none_rows(self, K):
        """Sorted rows whose suffix is shorter than K (get_ref_seq returns None there, RMI_LUT.py:89-92)."""
        return np.sort(np.nonzero(self.sa.astype(np.int64) - 1 + K > int(self.info.n_bases))[0]).astype(np.uint32)

    def locate(self, rows, sample):
        rows = np.ascontiguousarray(rows, np.uint32)
        out = np.zeros(len(rows), np.uint32)
        self.lib.emu_locate(C.byref(self.e), sample, len(rows), rows.ctypes.data, out.ctypes.data)
        return out

    def pack_read(self, q):
        lens = np.asarray([len(q)], np.uint32)
        off = np.zeros(2, np.uint32)
        self.capi.check(self.capi.lib.gsm_pack_reads(q.encode(), lens.ctypes.data, 1, off.ctypes.data, None))
        buf = np.zeros(int(off[1]) * 4 + 4, np.uint32)
        self.capi.check(self.capi.lib.gsm_pack_reads(q.encode(), lens.ctypes.data, 1, off.ctypes.data, buf.ctypes.data))
        return buf

    def sweep(self, q):
        w = self.pack_read(q)
        out = np.zeros(4 * (len(q) + 1), np.uint32)
        steps = C.c_uint64()
        sk, st = self._seed_args()
        n = self.lib.emu_sweep(C.byref(self.e), w.ctypes.data, len(q), out.ctypes.data, len(q) + 1, C.byref(steps), sk, st)
        return [tuple(int(x) for x in out[4 * k:4 * k + 4]) for k in range(n)], steps.value

    def bwa_picks(self, mems):
        """Pick mask of the sweep's hand-over (bwa_pick_key) for an ordered match list [(start, end, ...), ...] of <= 32 entries."""
        st = np.asarray([m[0] for m in mems], np.uint32)
        en = np.asarray([m[1] for m in mems], np.uint32)
        return int(self.lib.emu_bwa_picks(len(mems), st.ctypes.data, en.ctypes.data))

    def lut(self, K):
        if K not in self._lut:
            t = np.zeros(2 << (2 * K), np.uint32)
            self.lib.emu_lut_build(C.byref(self.e), K, t.ctypes.data)
            self._lut[K] = t
        return self._lut[K]

    def smem(self, method, q, min_len=1, K=0, rmi=None):
        """-> list of (i, j, lo, hi) records, or 'raises' / 'short'."""
        w = self.pack_read(q)
        cap = len(q) + 1
        out = np.zeros(6 * cap, np.uint32)
        lut_p = self.lut(K).ctypes.data if method == 1 else None
        sk, st = self._seed_args()
        if method == 2:
            K = rmi["K"]
            ls = np.asarray(rmi["level_sizes"], np.uint32)
            coef = np.ascontiguousarray(rmi["coef"], np.float64)
            icpt = np.ascontiguousarray(rmi["intercept"], np.float64)
            nr = self.none_rows(K) if self.rmi_fast else np.zeros(0, np.uint32)
            n = self.lib.emu_smem(C.byref(self.e), 2, w.ctypes.data, len(q), min_len, K, None, len(ls), ls.ctypes.data,
                                  coef.ctypes.data, icpt.ctypes.data, out.ctypes.data, cap, sk, st, len(nr), nr.ctypes.data if len(nr) else None)
        else:
            n = self.lib.emu_smem(C.byref(self.e), method, w.ctypes.data, len(q), min_len, K, lut_p, 0, None, None, None,
                                  out.ctypes.data, cap, sk, st, 0, None)
        if n == -1:
            return "raises"
        if n == -2:
            return "short"
        o64 = out[:6 * n].reshape(n, 6).astype(np.int64)
        lo = (o64[:, 2] | (o64[:, 3] << 32)).astype(np.int64)
        hi = (o64[:, 4] | (o64[:, 5] << 32)).astype(np.int64)
        return [(int(o64[k, 0]), int(o64[k, 1]), int(lo[k]), int(hi[k])) for k in range(n)]

    def rmi_lookup(self, rmi, code):
        ls = np.asarray(rmi["level_sizes"], np.uint32)
        coef = np.ascontiguousarray(rmi["coef"], np.float64)
        icpt = np.ascontiguousarray(rmi["intercept"], np.float64)
        pred, lo, hi = C.c_double(), C.c_int64(), C.c_int64()
        st = self.lib.emu_rmi_lookup(C.byref(self.e), rmi["K"], len(ls), ls.ctypes.data, coef.ctypes.data, icpt.ctypes.data,
                                     C.c_uint64(code), C.byref(pred), C.byref(lo), C.byref(hi))
        return st, pred.value, lo.value, hi.value


def _rmi_search(self, rmi, code):
    """(status, lo, hi, n_probes) through RmiSearch, the single-probe-site machine the kernels run."""
    ls = np.asarray(rmi["level_sizes"], np.uint32)
    coef = np.ascontiguousarray(rmi["coef"], np.float64)
    icpt = np.ascontiguousarray(rmi["intercept"], np.float64)
    lo, hi, npb = C.c_int64(), C.c_int64(), C.c_uint32()
    st = self.lib.emu_rmi_search(C.byref(self.e), rmi["K"], len(ls), ls.ctypes.data, coef.ctypes.data, icpt.ctypes.data,
                                 C.c_uint64(code), C.byref(lo), C.byref(hi), C.byref(npb))
    return st, lo.value, hi.value, npb.value


Emu.rmi_search = _rmi_search


def _rmi_fast(self, rmi, code):
    """(hazard, lo, hi, n_probes) through rmi_fast_lookup, the error-bounded search of the common case."""
    ls = np.asarray(rmi["level_sizes"], np.uint32)
    coef = np.ascontiguousarray(rmi["coef"], np.float64)
    icpt = np.ascontiguousarray(rmi["intercept"], np.float64)
    nr = self.none_rows(rmi["K"])
    lo, hi, hz, npb = C.c_int64(), C.c_int64(), C.c_uint32(), C.c_uint32()
    self.lib.emu_rmi_fast(C.byref(self.e), rmi["K"], len(ls), ls.ctypes.data, coef.ctypes.data, icpt.ctypes.data, len(nr), nr.ctypes.data,
                          C.c_uint64(code), C.byref(lo), C.byref(hi), C.byref(hz), C.byref(npb))
    return bool(hz.value), lo.value, hi.value, npb.value


Emu.rmi_fast_lookup = _rmi_fast


def _rmi_arith(self, rmi, code):
    """(hazard, lo, hi) through rmi_arith_lookup: true bounds from the FM index + arithmetic replay, no table probes."""
    ls = np.asarray(rmi["level_sizes"], np.uint32)
    coef = np.ascontiguousarray(rmi["coef"], np.float64)
    icpt = np.ascontiguousarray(rmi["intercept"], np.float64)
    nr = self.none_rows(rmi["K"])
    lo, hi, hz = C.c_int64(), C.c_int64(), C.c_uint32()
    self.lib.emu_rmi_arith(C.byref(self.e), rmi["K"], len(ls), ls.ctypes.data, coef.ctypes.data, icpt.ctypes.data, len(nr), nr.ctypes.data,
                           C.c_uint64(code), C.byref(lo), C.byref(hi), C.byref(hz))
    return bool(hz.value), lo.value, hi.value


Emu.rmi_arith_lookup = _rmi_arith


def _rmi_hazards(self, rmi):
    """Sorted hazard codes of a model on this index (rmi_code_is_hazard over all 4^K codes): the set k_rmi_hazard_scan finds."""
    ls = np.asarray(rmi["level_sizes"], np.uint32)
    coef = np.ascontiguousarray(rmi["coef"], np.float64)
    icpt = np.ascontiguousarray(rmi["intercept"], np.float64)
    nr = self.none_rows(rmi["K"])
    cap = 1 << (2 * rmi["K"])
    out = np.zeros(cap, np.uint32)
    n = self.lib.emu_rmi_hazards(C.byref(self.e), rmi["K"], len(ls), ls.ctypes.data, coef.ctypes.data, icpt.ctypes.data, len(nr), nr.ctypes.data,
                                 out.ctypes.data, cap)
    return np.sort(out[:n])


Emu.rmi_hazards = _rmi_hazards


def hazard_table(lib, codes):
    """Open-addressing table of the codes (hz_build), sized as RmiParams.build_hazard_filter sizes it; None if it refuses."""
    codes = np.ascontiguousarray(codes, np.uint32)
    n_slots = 1024
    while n_slots < 4 * len(codes) + 4:
        n_slots *= 2
    slots = np.zeros(n_slots, np.uint32)
    return slots if lib.emu_hz_build(codes.ctypes.data, len(codes), slots.ctypes.data, n_slots) else None


def records_to_dict(q, recs):
    """What the reference's result dict holds for these records (duplicate strings collapse,
    first insertion keeps its place, last value wins) as [[key, lo, hi], ...]."""
    d = {}
    for i, j, lo, hi in recs:
        d[q[i:j]] = (lo, hi)
    return [[k, v[0], v[1]] for k, v in d.items()]
