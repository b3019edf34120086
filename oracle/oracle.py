"""ctypes binding of oracle/skani_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this.
The product package (skder_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libskani_oracle.so")


def build(force=False):
    src = [os.path.join(_HERE, "skani_oracle.c"), os.path.join(_HERE, "skani_oracle.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(
        ["gcc", "-O2", "-std=c99", "-D_GNU_SOURCE", "-pthread", "-shared", "-fPIC", "-o", _SO, src[0], "-lz", "-lm"]
    )
    return _SO


class Params(C.Structure):
    _fields_ = [
        ("k", C.c_int32),
        ("marker_k", C.c_int32),
        ("c", C.c_uint64),
        ("marker_c", C.c_uint64),
        ("min_contig_len", C.c_int32),
        ("contig_pad", C.c_int32),
        ("chunk_len", C.c_int32),
        ("band_bp", C.c_int32),
        ("max_gap", C.c_int32),
        ("lookback", C.c_int32),
        ("anchor_score", C.c_int32),
        ("min_anchors", C.c_int32),
        ("min_score", C.c_int32),
        ("max_mult", C.c_int32),
        ("max_chunk_anchors", C.c_int32),
        ("max_chunk_chains", C.c_int32),
        ("ovl_num", C.c_int32),
        ("ovl_den", C.c_int32),
        ("span_ext", C.c_int32),
        ("role_rule", C.c_int32),
    ]


class Chain(C.Structure):
    _fields_ = [
        ("chunk", C.c_int32),
        ("n_anchors", C.c_int32),
        ("n_seeds", C.c_int32),
        ("score", C.c_int32),
        ("q0", C.c_uint32),
        ("q1", C.c_uint32),
        ("r0", C.c_uint32),
        ("r1", C.c_uint32),
        ("rev", C.c_int32),
    ]


class PairResult(C.Structure):
    _fields_ = [
        ("ani", C.c_double),
        ("ani_raw", C.c_double),
        ("af_a", C.c_double),
        ("af_b", C.c_double),
        ("n_chains", C.c_int32),
        ("swapped", C.c_int32),
        ("n_anchors_total", C.c_int64),
        ("n_seeds_total", C.c_int64),
        ("span_q", C.c_int64),
        ("span_r", C.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ora_default_params.argtypes = [C.POINTER(Params)]
        L.ora_mm_hash64.restype = C.c_uint64
        L.ora_mm_hash64.argtypes = [C.c_uint64]
        L.ora_sketch_file.restype = C.c_void_p
        L.ora_sketch_file.argtypes = [C.c_char_p, C.POINTER(Params)]
        L.ora_sketch_contigs.restype = C.c_void_p
        L.ora_sketch_contigs.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int, C.POINTER(Params)]
        L.ora_sketch_free.argtypes = [C.c_void_p]
        for name, rt in [
            ("ora_n_seeds", C.c_int64),
            ("ora_n_markers", C.c_int64),
            ("ora_total_len", C.c_int64),
            ("ora_n_contigs", C.c_int32),
            ("ora_n_chunks", C.c_int32),
            ("ora_seeds", C.POINTER(C.c_uint64)),
            ("ora_markers", C.POINTER(C.c_uint64)),
            ("ora_contig_lens", C.POINTER(C.c_int64)),
            ("ora_first_name", C.c_char_p),
        ]:
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [C.c_void_p]
        L.ora_screen.restype = C.c_int64
        L.ora_screen.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.POINTER(Params), C.POINTER(C.c_int)]
        L.ora_pair.restype = C.c_int
        L.ora_pair.argtypes = [
            C.c_void_p,
            C.c_void_p,
            C.POINTER(Params),
            C.POINTER(PairResult),
            C.POINTER(Chain),
            C.c_int,
            C.POINTER(C.c_int),
        ]
        _lib = L
    return _lib


def default_params():
    p = Params()
    lib().ora_default_params(C.byref(p))
    return p


class Sketch:
    def __init__(self, handle, path=None):
        self.h = handle
        self.path = path

    @classmethod
    def from_file(cls, path, params=None):
        p = params or default_params()
        h = lib().ora_sketch_file(os.fsencode(path), C.byref(p))
        if not h:
            raise IOError("oracle: cannot read %s" % path)
        return cls(h, path)

    @classmethod
    def from_contigs(cls, seqs, params=None):
        p = params or default_params()
        n = len(seqs)
        arr = (C.c_char_p * n)(*[s if isinstance(s, bytes) else s.encode() for s in seqs])
        lens = (C.c_int64 * n)(*[len(s) for s in seqs])
        return cls(lib().ora_sketch_contigs(arr, lens, n, C.byref(p)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_sketch_free(self.h)
            self.h = None

    @property
    def n_seeds(self):
        return lib().ora_n_seeds(self.h)

    @property
    def n_markers(self):
        return lib().ora_n_markers(self.h)

    @property
    def total_len(self):
        return lib().ora_total_len(self.h)

    @property
    def n_contigs(self):
        return lib().ora_n_contigs(self.h)

    @property
    def n_chunks(self):
        return lib().ora_n_chunks(self.h)

    @property
    def first_name(self):
        return lib().ora_first_name(self.h).decode()

    def seeds(self):
        n = self.n_seeds
        return np.ctypeslib.as_array(lib().ora_seeds(self.h), shape=(n,)).copy() if n else np.zeros(0, np.uint64)

    def markers(self):
        n = self.n_markers
        return np.ctypeslib.as_array(lib().ora_markers(self.h), shape=(n,)).copy() if n else np.zeros(0, np.uint64)

    def contig_lens(self):
        n = self.n_contigs
        return np.ctypeslib.as_array(lib().ora_contig_lens(self.h), shape=(n,)).copy() if n else np.zeros(0, np.int64)


def screen(a, b, s, params=None):
    p = params or default_params()
    ok = C.c_int(0)
    shared = lib().ora_screen(a.h, b.h, float(s), C.byref(p), C.byref(ok))
    return shared, bool(ok.value)


def triangle(sketches, s, min_af, threads, params=None, want_pass=False):
    """All-vs-all of `sketches` on host threads inside the C library (no Python per pair): inverted-index prescreen,
    ANI/AF of the survivors.  Returns a dict: survivors, edges, t_index, t_count, t_ani (seconds) and, if want_pass,
    the n x n uint8 decision matrix (row a, column b, a < b)."""
    p = params or default_params()
    n = len(sketches)
    arr = (C.c_void_p * n)(*[sk.h for sk in sketches])
    ne, ti, tc, ta = C.c_int64(0), C.c_double(0), C.c_double(0), C.c_double(0)
    mat = np.zeros((n, n), np.uint8) if want_pass else None
    f = lib().ora_triangle
    f.restype = C.c_int64
    f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_void_p]
    ns = f(arr, n, float(s), float(min_af), C.byref(p), int(threads), C.byref(ne),
           mat.ctypes.data if want_pass else None, C.byref(ti), C.byref(tc), C.byref(ta))
    if ns < 0:
        raise ValueError("oracle triangle: too many genomes")
    out = {"survivors": int(ns), "edges": int(ne.value), "t_index": ti.value, "t_count": tc.value, "t_ani": ta.value}
    if want_pass:
        out["pass"] = mat
    return out


def search(sketches, q, s, min_af, threads, params=None):
    """One query (index q into `sketches`) against all of them inside the C library: [(ref index, ani, af_ref, af_query)]
    sorted by ref index; fractions in [0, 1]."""
    p = params or default_params()
    n = len(sketches)
    arr = (C.c_void_p * n)(*[sk.h for sk in sketches])
    ref = np.zeros(n, np.int32)
    ani, afr, afq = np.zeros(n), np.zeros(n), np.zeros(n)
    f = lib().ora_search
    f.restype = C.c_int64
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p]
    k = f(arr, n, int(q), float(s), float(min_af), C.byref(p), int(threads), ref.ctypes.data, ani.ctypes.data, afr.ctypes.data,
          afq.ctypes.data)
    o = np.argsort(ref[:k])
    return [(int(ref[i]), float(ani[i]), float(afr[i]), float(afq[i])) for i in o]


def pair(a, b, params=None, want_chains=False, max_chains=8192):
    p = params or default_params()
    r = PairResult()
    if want_chains:
        ch = (Chain * max_chains)()
        n = C.c_int(0)
        lib().ora_pair(a.h, b.h, C.byref(p), C.byref(r), ch, max_chains, C.byref(n))
        return r, ch[: n.value]
    lib().ora_pair(a.h, b.h, C.byref(p), C.byref(r), None, 0, None)
    return r
