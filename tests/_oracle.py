"""ctypes view of the CPU oracle (oracle/mash_screen_oracle.c) for the tests.

TEST INFRASTRUCTURE.  Builds oracle/_build/liboracle.so with `make -C oracle`
when missing (gcc only).  The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
BIN = os.path.join(ROOT, "oracle", "_build", "oracle_mash")


def build() -> None:
    src = os.path.join(ROOT, "oracle", "mash_screen_oracle.c")
    if (not os.path.exists(LIB) or not os.path.exists(BIN)
            or os.path.getmtime(LIB) < os.path.getmtime(src)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True,
                       stdout=subprocess.DEVNULL)


class OrcStats(C.Structure):
    _fields_ = [("set_size", C.c_uint64), ("n_bases", C.c_uint64), ("n_records", C.c_uint64),
                ("n_kmers", C.c_uint64), ("n_mixture", C.c_uint64),
                ("t_stream", C.c_double), ("t_reduce", C.c_double)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        u64p, u32p, u8p, f64p = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                 C.POINTER(C.c_uint8), C.POINTER(C.c_double))
        L.orc_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, u64p]
        L.orc_murmur3_x64_128.restype = None
        L.orc_use64.argtypes = [C.c_uint32]; L.orc_use64.restype = C.c_int
        L.orc_hash_sequence.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, u64p, u8p]
        L.orc_hash_sequence.restype = C.c_int
        L.orc_db_from_arrays.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, u64p, u64p, u64p]
        L.orc_db_from_arrays.restype = C.c_void_p
        L.orc_db_load_msh.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.orc_db_load_msh.restype = C.c_void_p
        L.orc_db_free.argtypes = [C.c_void_p]; L.orc_db_free.restype = None
        for f in ("n_refs", "n_entries", "n_distinct"):
            getattr(L, "orc_db_" + f).argtypes = [C.c_void_p]
            getattr(L, "orc_db_" + f).restype = C.c_uint64
        for f in ("k", "s", "seed"):
            getattr(L, "orc_db_" + f).argtypes = [C.c_void_p]
            getattr(L, "orc_db_" + f).restype = C.c_uint32
        for f in ("name", "comment"):
            getattr(L, "orc_db_" + f).argtypes = [C.c_void_p, C.c_uint64]
            getattr(L, "orc_db_" + f).restype = C.c_char_p
        for f in ("size", "length"):
            getattr(L, "orc_db_" + f).argtypes = [C.c_void_p, C.c_uint64]
            getattr(L, "orc_db_" + f).restype = C.c_uint64
        L.orc_db_hashes.argtypes = [C.c_void_p, C.c_uint64]; L.orc_db_hashes.restype = u64p
        L.orc_identity.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]; L.orc_identity.restype = C.c_double
        L.orc_pvalue.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_uint64]; L.orc_pvalue.restype = C.c_double
        L.orc_set_size.argtypes = [u64p, C.c_uint64, C.c_int]; L.orc_set_size.restype = C.c_uint64
        L.orc_sketch_text.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                      u64p, u64p, u64p]
        L.orc_sketch_text.restype = C.c_int
        L.orc_screen_text.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_int,
                                      u64p, u32p, f64p, f64p, u32p, u64p, C.POINTER(OrcStats)]
        L.orc_screen_text.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def murmur(data: bytes, seed: int = 42):
    out = (C.c_uint64 * 2)()
    lib().orc_murmur3_x64_128(data, len(data), seed, out)
    return int(out[0]), int(out[1])


def hash_sequence(seq: bytes, k: int, seed: int = 42):
    n = len(seq)
    m = max(n - k + 1, 0)
    h = np.zeros(m, np.uint64); v = np.zeros(m, np.uint8)
    if m:
        rc = lib().orc_hash_sequence(seq, n, k, seed, _p(h, C.c_uint64), _p(v, C.c_uint8))
        assert rc == 0
    return h, v.astype(bool)


def sketch_text(text: bytes, k: int, s: int, seed: int = 42, threads: int = 1):
    out = np.zeros(s + 1, np.uint64)
    n = C.c_uint64(0); ln = C.c_uint64(0)
    rc = lib().orc_sketch_text(text, len(text), k, s, seed, threads, _p(out, C.c_uint64),
                               C.byref(n), C.byref(ln))
    assert rc == 0
    return out[:n.value].copy(), int(ln.value)


@dataclass
class ScreenResult:
    shared: np.ndarray
    median: np.ndarray
    identity: np.ndarray
    pvalue: np.ndarray
    counts_per_entry: np.ndarray
    mixture: np.ndarray
    set_size: int
    n_bases: int
    n_kmers: int
    t_stream: float
    t_reduce: float


class OracleDB:
    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_arrays(cls, k, s, seed, offsets, hashes, lengths):
        offsets = np.ascontiguousarray(offsets, np.uint64)
        hashes = np.ascontiguousarray(hashes, np.uint64)
        lengths = np.ascontiguousarray(lengths, np.uint64)
        h = lib().orc_db_from_arrays(k, s, seed, len(offsets) - 1, _p(offsets, C.c_uint64),
                                     _p(hashes, C.c_uint64), _p(lengths, C.c_uint64))
        assert h
        return cls(h)

    @classmethod
    def load_msh(cls, path: str):
        err = C.create_string_buffer(512)
        h = lib().orc_db_load_msh(path.encode(), err, 512)
        if not h:
            raise RuntimeError(err.value.decode())
        return cls(h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_db_free(self.h)
            self.h = None

    @property
    def n_refs(self): return int(lib().orc_db_n_refs(self.h))
    @property
    def n_entries(self): return int(lib().orc_db_n_entries(self.h))
    @property
    def n_distinct(self): return int(lib().orc_db_n_distinct(self.h))
    @property
    def k(self): return int(lib().orc_db_k(self.h))
    @property
    def s(self): return int(lib().orc_db_s(self.h))
    @property
    def seed(self): return int(lib().orc_db_seed(self.h))
    def name(self, i): return lib().orc_db_name(self.h, i).decode()
    def comment(self, i): return lib().orc_db_comment(self.h, i).decode()
    def size(self, i): return int(lib().orc_db_size(self.h, i))
    def length(self, i): return int(lib().orc_db_length(self.h, i))

    def hashes(self, i):
        n = self.size(i)
        p = lib().orc_db_hashes(self.h, i)
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint64)

    def screen_text(self, text: bytes, threads: int = 1, wta: bool = False) -> ScreenResult:
        N, E, s = self.n_refs, self.n_entries, self.s
        shared = np.zeros(N, np.uint64); median = np.zeros(N, np.uint32)
        ident = np.zeros(N, np.float64); pv = np.zeros(N, np.float64)
        cpe = np.zeros(E + 1, np.uint32); mix = np.zeros(s + 1, np.uint64)
        st = OrcStats()
        rc = lib().orc_screen_text(self.h, text, len(text), threads, int(wta),
                                   _p(shared, C.c_uint64), _p(median, C.c_uint32),
                                   _p(ident, C.c_double), _p(pv, C.c_double),
                                   _p(cpe, C.c_uint32), _p(mix, C.c_uint64), C.byref(st))
        assert rc == 0
        return ScreenResult(shared, median, ident, pv, cpe[:E], mix[:st.n_mixture].copy(),
                            int(st.set_size), int(st.n_bases), int(st.n_kmers),
                            float(st.t_stream), float(st.t_reduce))


def fmt_g(x: float) -> str:
    """C's %g with precision 6 == C++ ostream default (S16)."""
    return "%g" % x
