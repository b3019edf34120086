"""Host-side mirror of the skani sub-commands skDER uses, on top of the C-ABI (include/skani_b200.h).

  skani sketch   -> Engine.add_fasta() + Engine.save()      (reference src/skDER/skder.py:103)
  skani triangle -> Engine.triangle()                       (skder.py:16-18)
  skani dist     -> Engine.rect(refs, queries)              (skder.py:58-59, cidder.py:362-363)
  skani search   -> Engine.load() + add query + rect()      (skder.py:119)

All compute happens in libskani_b200.so on the GPU; nothing here estimates anything.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import SkbError

EDGE_DTYPE = np.dtype(
    {"names": ["a", "b", "ani", "af_a", "af_b"], "formats": ["<u4", "<u4", "<f8", "<f8", "<f8"],
     "offsets": [0, 4, 8, 16, 24], "itemsize": 32}
)
assert C.sizeof(_lib.Edge) == EDGE_DTYPE.itemsize


def default_params():
    p = _lib.Params()
    _lib.lib().skb_default_params(C.byref(p))
    return p


class PackedGenome:
    """2-bit packed genome in host memory (owned by the library)."""

    def __init__(self, ptr, path=None):
        self._p = ptr
        self.path = path

    def __del__(self):
        if getattr(self, "_p", None):
            _lib.lib().skb_packed_free(self._p)
            self._p = None

    n_bases = property(lambda s: s._p.contents.n_bases)
    n_words = property(lambda s: s._p.contents.n_words)
    n_contigs = property(lambda s: s._p.contents.n_contigs)
    n50 = property(lambda s: s._p.contents.n50)
    total_bases_all = property(lambda s: s._p.contents.total_bases_all)

    @property
    def first_name(self):
        return (self._p.contents.first_name or b"").decode("utf-8", "replace")

    def contig_lens(self):
        n = self.n_contigs
        return np.ctypeslib.as_array(self._p.contents.contig_lens, shape=(n,)).copy() if n else np.zeros(0, np.int64)

    def words(self):
        return np.ctypeslib.as_array(self._p.contents.words, shape=(self.n_words,))


def pack_fasta(path, min_contig_len=500):
    out = C.POINTER(_lib.Packed)()
    rc = _lib.lib().skb_pack_fasta(os.fsencode(path), min_contig_len, C.byref(out))
    if rc != 0 or not out:
        raise SkbError("cannot read FASTA %s (code %d)" % (path, rc))
    return PackedGenome(out, path)


def pack_fasta_many(paths, threads=None, min_contig_len=500):
    n = len(paths)
    if n == 0:
        return []
    arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
    out = (C.POINTER(_lib.Packed) * n)()
    threads = threads or os.cpu_count() or 1
    nfail = _lib.lib().skb_pack_fasta_many(arr, n, min_contig_len, threads, out)
    res = [PackedGenome(out[i], paths[i]) if out[i] else None for i in range(n)]
    if nfail:
        bad = [paths[i] for i in range(n) if res[i] is None]
        raise SkbError("cannot read %d FASTA file(s), first: %s" % (len(bad), bad[0]))
    return res


def pack_contigs(seqs, min_contig_len=500):
    seqs = [s if isinstance(s, bytes) else s.encode() for s in seqs]
    n = len(seqs)
    arr = (C.c_char_p * n)(*seqs)
    lens = (C.c_int64 * n)(*[len(s) for s in seqs])
    out = C.POINTER(_lib.Packed)()
    rc = _lib.lib().skb_pack_contigs(arr, lens, n, min_contig_len, C.byref(out))
    if rc != 0:
        raise SkbError("pack_contigs failed (code %d)" % rc)
    return PackedGenome(out)


class Engine:
    def __init__(self, device=0, params=None):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.params = params or default_params()
        rc = self._L.skb_create(int(device), C.byref(self.params), C.byref(self._h))
        if rc != 0:
            raise SkbError("skb_create failed (%d): %s" % (rc, self._L.skb_last_error(None).decode()))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.skb_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc, what):
        if rc != 0:
            raise SkbError("%s failed (%d): %s" % (what, rc, self._L.skb_last_error(self._h).decode()))

    # ---- sketching -------------------------------------------------------------------------
    def add(self, packed):
        n = len(packed)
        if n == 0:
            return
        arr = (C.POINTER(_lib.Packed) * n)(*[p._p for p in packed])
        self._ck(self._L.skb_add_genomes(self._h, n, arr), "skb_add_genomes")

    def add_fasta(self, paths, threads=None):
        packed = pack_fasta_many(list(paths), threads, self.params.min_contig_len)
        self.add(packed)
        return packed

    def clear(self, keep_tables=False):
        if keep_tables:
            self._ck(self._L.skb_clear_keep_tables(self._h), "skb_clear_keep_tables")
        else:
            self._ck(self._L.skb_clear(self._h), "skb_clear")

    def index_seed_tables(self):
        self._ck(self._L.skb_index_seed_tables(self._h), "skb_index_seed_tables")

    def timer_start(self):
        self._ck(self._L.skb_timer_start(self._h), "skb_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self._L.skb_timer_stop(self._h, C.byref(ms)), "skb_timer_stop")
        return ms.value

    def index(self):
        self._ck(self._L.skb_index(self._h), "skb_index")

    def index_append(self):
        self._ck(self._L.skb_index_append(self._h), "skb_index_append")

    def pop_last_add(self):
        self._ck(self._L.skb_pop_last_add(self._h), "skb_pop_last_add")

    def search(self, query_packed, screen=80.0, min_af=15.0):
        """`skani search`: one query genome against the resident, indexed database (ids 0..n-1).  The query
        is sketched, indexed query-only, compared, and removed again.  Edges: a = database genome, b = query."""
        n_db = self.n_genomes
        self.add([query_packed])
        try:
            self.index_append()
            return self.rect(np.arange(n_db, dtype=np.int32), [n_db], screen=screen, min_af=min_af)
        finally:
            self.pop_last_add()

    @property
    def n_genomes(self):
        return self._L.skb_n_genomes(self._h)

    @property
    def launches(self):
        return self._L.skb_launch_count(self._h)

    @property
    def stream(self):
        return self._L.skb_stream(self._h)

    def sizes(self, g):
        ns, nm, tl = C.c_int64(), C.c_int64(), C.c_int64()
        nc = C.c_int32()
        self._ck(self._L.skb_sketch_sizes(self._h, g, C.byref(ns), C.byref(nm), C.byref(nc), C.byref(tl)), "skb_sketch_sizes")
        return {"n_seeds": ns.value, "n_markers": nm.value, "n_chunks": nc.value, "total_len": tl.value}

    def seeds(self, g):
        ns = C.c_int64()
        self._ck(self._L.skb_sketch_sizes(self._h, g, C.byref(ns), None, None, None), "skb_sketch_sizes")
        out = np.zeros(ns.value, np.uint64)
        self._ck(self._L.skb_get_seeds(self._h, g, out.ctypes.data), "skb_get_seeds")
        return out

    def markers(self, g):
        out = np.zeros(self.sizes(g)["n_markers"], np.uint64)
        self._ck(self._L.skb_get_markers(self._h, g, out.ctypes.data), "skb_get_markers")
        return out

    def save(self, directory):
        os.makedirs(directory, exist_ok=True)
        self._ck(self._L.skb_db_save(self._h, os.fsencode(directory)), "skb_db_save")

    def load(self, directory):
        self._ck(self._L.skb_db_load(self._h, os.fsencode(directory)), "skb_db_load")

    # ---- pairs -----------------------------------------------------------------------------
    def _take_edges(self, ptr, n):
        try:
            if n.value == 0:
                return np.zeros(0, EDGE_DTYPE)
            buf = C.cast(ptr, C.POINTER(C.c_char * (n.value * EDGE_DTYPE.itemsize))).contents
            return np.frombuffer(buf, dtype=EDGE_DTYPE, count=n.value).copy()
        finally:
            self._L.skb_free(ptr)

    def triangle(self, screen=80.0, min_af=15.0, part=0, n_parts=1, to_host=True):
        """to_host=False leaves the edges on the device (see device_edges) and returns (None, stats)."""
        ptr, n, st = C.POINTER(_lib.Edge)(), C.c_int64(), _lib.Stats()
        self._ck(
            self._L.skb_triangle(self._h, float(screen), float(min_af), part, n_parts,
                                 C.byref(ptr) if to_host else None, C.byref(n), C.byref(st)),
            "skb_triangle",
        )
        return (self._take_edges(ptr, n) if to_host else None), st

    def set_owned(self, first, count):
        """Seed tables are built (next index()) only for genomes first .. first+count-1; count = -1: all."""
        self._ck(self._L.skb_set_owned(self._h, int(first), int(count)), "skb_set_owned")

    def screen_triangle(self, screen=80.0, part=0, n_parts=1):
        """Prescreen of this partition's rows; returns (device pointer to the sorted (a << 32 | b) pairs, count, stats)."""
        ptr, n, st = C.c_void_p(), C.c_int64(), _lib.Stats()
        self._ck(self._L.skb_screen_triangle(self._h, float(screen), part, n_parts, C.byref(ptr), C.byref(n), C.byref(st)),
                 "skb_screen_triangle")
        return int(ptr.value or 0), int(n.value), st

    def pairs_edges(self, dev_pairs, n_pairs, owned_only=True, min_af=15.0, to_host=True):
        """ANI/AF of a device-resident pair list (only the pairs whose reference this context owns, if owned_only)."""
        ptr, n, st = C.POINTER(_lib.Edge)(), C.c_int64(), _lib.Stats()
        self._ck(
            self._L.skb_pairs_edges(self._h, C.c_void_p(dev_pairs), int(n_pairs), 1 if owned_only else 0, float(min_af),
                                    C.byref(ptr) if to_host else None, C.byref(n), C.byref(st)),
            "skb_pairs_edges",
        )
        return (self._take_edges(ptr, n) if to_host else None), st

    def greedy_summary(self, edges, n_genomes, min_ani, min_af):
        """(connectivity[n], member_off[n+1], members[]) of a binary edge list -- reference skDERsum's first pass.
        edges: EDGE_DTYPE array, or None for the list the last triangle/rect left on the device."""
        conn, off, mem = C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_uint32)()
        if edges is not None:
            edges = np.ascontiguousarray(edges, EDGE_DTYPE)
        self._ck(
            self._L.skb_greedy_summary(self._h, edges.ctypes.data if edges is not None else None,
                                       len(edges) if edges is not None else 0, int(n_genomes), float(min_ani), float(min_af),
                                       C.byref(conn), C.byref(off), C.byref(mem)),
            "skb_greedy_summary",
        )
        try:
            c = np.ctypeslib.as_array(conn, shape=(max(n_genomes, 1),))[:n_genomes].copy()
            o = np.ctypeslib.as_array(off, shape=(n_genomes + 1,)).copy()
            m = np.ctypeslib.as_array(mem, shape=(max(int(o[-1]), 1),))[: int(o[-1])].copy()
        finally:
            for p in (conn, off, mem):
                self._L.skb_free(p)
        return c, o, m

    def device_edges(self):
        """(device pointer, count) of the last triangle/rect result; valid until the next call on this engine."""
        ptr, n = C.c_void_p(), C.c_int64()
        self._ck(self._L.skb_device_edges(self._h, C.byref(ptr), C.byref(n)), "skb_device_edges")
        return int(ptr.value or 0), int(n.value)

    def rect(self, refs, queries, screen=80.0, min_af=15.0):
        refs = np.ascontiguousarray(refs, np.int32)
        queries = np.ascontiguousarray(queries, np.int32)
        ptr, n, st = C.POINTER(_lib.Edge)(), C.c_int64(), _lib.Stats()
        self._ck(
            self._L.skb_rect(self._h, refs.ctypes.data, len(refs), queries.ctypes.data, len(queries), float(screen),
                             float(min_af), C.byref(ptr), C.byref(n), C.byref(st)),
            "skb_rect",
        )
        return self._take_edges(ptr, n), st

    def pairs_detail(self, a, b):
        a = np.ascontiguousarray(a, np.uint32)
        b = np.ascontiguousarray(b, np.uint32)
        out = (_lib.PairDetail * len(a))()
        self._ck(self._L.skb_pairs_detail(self._h, a.ctypes.data, b.ctypes.data, len(a), out), "skb_pairs_detail")
        return out

    def shared_markers(self, a, b):
        a = np.ascontiguousarray(a, np.uint32)
        b = np.ascontiguousarray(b, np.uint32)
        out = np.zeros(len(a), np.int64)
        self._ck(self._L.skb_shared_markers(self._h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data), "skb_shared_markers")
        return out

    def sketch_view(self):
        v = _lib.SketchView()
        self._ck(self._L.skb_sketch_view_get(self._h, C.byref(v)), "skb_sketch_view_get")
        return v
