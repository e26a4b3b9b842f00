"""ctypes wrapper of oracle/smem_oracle.c -- TEST INFRASTRUCTURE (see the header of that file).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsmem_oracle.so")


def build():
    src = os.path.join(HERE, "smem_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB


class COracle:
    """index = text + 1-based suffix array (the reference's fm_index arrays)."""

    def __init__(self, text, suffix_array_1based):
        self.lib = C.CDLL(build())
        P, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        self.lib.orc_create.restype = P
        self.lib.orc_create.argtypes = [C.c_char_p, u64, P]
        self.lib.orc_destroy.argtypes = [P]
        self.lib.orc_backsearch.argtypes = [P, C.c_char_p, P, u64, P, P]
        self.lib.orc_build_lut.argtypes = [P, i32]
        self.lib.orc_rmi_lookup.argtypes = [P, i32, i32, P, P, P, u64, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        self.lib.orc_smems.argtypes = [P, i32, C.c_char_p, P, u64, i32, i32, i32, P, P, P, P, u32, P, i32]
        self.text = text if isinstance(text, bytes) else text.encode()
        self.sa = np.ascontiguousarray(suffix_array_1based, dtype=np.uint32)
        self.h = self.lib.orc_create(self.text, len(self.text), self.sa.ctypes.data)
        self.max_threads = self.lib.orc_max_threads()

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    def backsearch(self, reads):
        lens = np.asarray([len(r) for r in reads], np.uint32)
        lo = np.zeros(len(reads), np.int64)
        hi = np.zeros(len(reads), np.int64)
        self.lib.orc_backsearch(self.h, "".join(reads).encode(), lens.ctypes.data, len(reads), lo.ctypes.data, hi.ctypes.data)
        return lo, hi

    def rmi_lookup(self, rmi, code):
        ls = np.asarray(rmi["level_sizes"], np.uint32)
        coef = np.ascontiguousarray(rmi["coef"], np.float64)
        icpt = np.ascontiguousarray(rmi["intercept"], np.float64)
        pred, lo, hi = C.c_double(), C.c_int64(), C.c_int64()
        st = self.lib.orc_rmi_lookup(self.h, rmi["K"], len(ls), ls.ctypes.data, coef.ctypes.data, icpt.ctypes.data, code,
                                     C.byref(pred), C.byref(lo), C.byref(hi))
        return st, pred.value, lo.value, hi.value

    def smems(self, method, reads, min_len=1, K=0, rmi=None, threads=0, joined=None, lens=None):
        """-> (out int64[n, cap, 4], counts int32[n]); counts -1 = reference raises, -2 = too short."""
        if joined is None:
            lens = np.asarray([len(r) for r in reads], np.uint32)
            joined = "".join(reads).encode()
        n = len(lens)
        cap = int(lens.max()) + 1 if n else 1
        out = np.zeros((n, cap, 4), np.int64)
        counts = np.zeros(n, np.int32)
        if method == 2:
            ls = np.asarray(rmi["level_sizes"], np.uint32)
            coef = np.ascontiguousarray(rmi["coef"], np.float64)
            icpt = np.ascontiguousarray(rmi["intercept"], np.float64)
            st = self.lib.orc_smems(self.h, 2, joined, lens.ctypes.data, n, min_len, rmi["K"], len(ls), ls.ctypes.data,
                                    coef.ctypes.data, icpt.ctypes.data, out.ctypes.data, cap, counts.ctypes.data, threads)
        else:
            st = self.lib.orc_smems(self.h, method, joined, lens.ctypes.data, n, min_len, K, 0, None, None, None,
                                    out.ctypes.data, cap, counts.ctypes.data, threads)
        if st != 0:
            raise ValueError("orc_smems: bad input")
        return out, counts

    def smem_dicts(self, method, reads, **kw):
        """[[key, lo, hi], ...] per read in the reference dict's order; 'raises' / 'short' markers."""
        out, counts = self.smems(method, reads, **kw)
        res = []
        for r, q in enumerate(reads):
            if counts[r] == -1:
                res.append("raises")
            elif counts[r] == -2:
                res.append("short")
            else:
                res.append([[q[int(o[0]):int(o[1])], int(o[2]), int(o[3])] for o in out[r, :counts[r]]])
        return res
