"""ctypes binding of oracle/libsw_oracle.so -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsw_oracle.so")


class OracleParams(C.Structure):
    _fields_ = [("match", C.c_int), ("mismatch", C.c_int), ("gap_open", C.c_int),
                ("gap_extend", C.c_int), ("score_width", C.c_int), ("first_col_v03", C.c_int)]


def build_oracle(force=False):
    """Compile the C oracle in place (gcc + OpenMP)."""
    src = os.path.join(_HERE, "sw_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsw_oracle.so"])
    return _LIB


class Oracle:
    def __init__(self, match=5, mismatch=-4, gap_open=-12, gap_extend=-4, score_width=0,
                 first_col_v03=0):
        if not os.path.exists(_LIB):
            build_oracle()
        self.lib = C.CDLL(_LIB)
        self.p = OracleParams(match, mismatch, gap_open, gap_extend, score_width, first_col_v03)
        L = self.lib
        L.swo_score_ascii.restype = C.c_int32
        L.swo_score_ascii.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(OracleParams)]
        L.swo_score_batch_packed.restype = C.c_int
        L.swo_score_batch_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                             C.POINTER(OracleParams), C.c_void_p, C.c_int]
        L.swo_max_threads.restype = C.c_int

    def score(self, q, t):
        qb, tb = q.encode(), t.encode()
        return int(self.lib.swo_score_ascii(qb, len(qb), tb, len(tb), C.byref(self.p)))

    def max_threads(self):
        return int(self.lib.swo_max_threads())

    def score_batch_packed(self, qpacked, qlen, qoff, tpacked, tlen, toff, nthreads=0):
        """All arrays numpy: packed uint8, len uint32, off uint64.  Returns (scores[nq, ns], threads)."""
        qpacked = np.ascontiguousarray(qpacked, dtype=np.uint8)
        tpacked = np.ascontiguousarray(tpacked, dtype=np.uint8)
        qlen = np.ascontiguousarray(qlen, dtype=np.uint32)
        tlen = np.ascontiguousarray(tlen, dtype=np.uint32)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        toff = np.ascontiguousarray(toff, dtype=np.uint64)
        nq, ns = len(qlen), len(tlen)
        out = np.zeros((nq, ns), dtype=np.int32)
        used = self.lib.swo_score_batch_packed(
            qpacked.ctypes.data, qlen.ctypes.data, qoff.ctypes.data, nq,
            tpacked.ctypes.data, tlen.ctypes.data, toff.ctypes.data, ns,
            C.byref(self.p), out.ctypes.data, nthreads)
        return out, int(used)
