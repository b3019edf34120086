"""CPU `skani triangle / dist / search` built on the C oracle -- TEST INFRASTRUCTURE ONLY.

Used by tests/ (to check the GPU shim's TSV byte-for-byte against the oracle's), by
__graft_entry__.smoke() and by bench.py's cpu_baseline / `--impl reference` legs.  It follows the
same flow skani does behind skDER's call sites (reference src/skDER/skder.py:16-18, :58-59, :119):
sketch every genome, screen pairs on markers, estimate ANI/AF for survivors, keep rows with
max(AF) >= --min-af, print 2-decimal TSV with Ref = lexicographically smaller path.
"""
import itertools
import os
from concurrent.futures import ThreadPoolExecutor

from . import oracle as O

HEADER = "Ref_file\tQuery_file\tANI\tAlign_fraction_ref\tAlign_fraction_query\tRef_name\tQuery_name\n"


def sketch_files(paths, threads=1, params=None):
    with ThreadPoolExecutor(max(1, threads)) as ex:
        return list(ex.map(lambda p: O.Sketch.from_file(p, params), paths))


def triangle_edges(sketches, screen_pct=80.0, min_af_pct=15.0, threads=1, params=None, pairs=None):
    """[(a, b, ani%, af_a%, af_b%)] for a<b, sorted."""
    n = len(sketches)
    pairs = list(itertools.combinations(range(n), 2)) if pairs is None else pairs

    def one(ab):
        a, b = ab
        if screen_pct > 0 and not O.screen(sketches[a], sketches[b], screen_pct / 100.0, params)[1]:
            return None
        r = O.pair(sketches[a], sketches[b], params)
        if r.ani < 0 or max(r.af_a, r.af_b) * 100.0 < min_af_pct:
            return None
        return (a, b, r.ani * 100.0, r.af_a * 100.0, r.af_b * 100.0)

    with ThreadPoolExecutor(max(1, threads)) as ex:
        res = list(ex.map(one, pairs, chunksize=64))
    return sorted(e for e in res if e is not None)


def triangle_tsv(paths, screen_pct=80.0, min_af_pct=15.0, threads=1):
    paths = sorted(set(paths))
    sk = sketch_files(paths, threads)
    names = [s.first_name for s in sk]
    rows = [
        "%s\t%s\t%.2f\t%.2f\t%.2f\t%s\t%s\n" % (paths[a], paths[b], ani, afa, afb, names[a], names[b])
        for a, b, ani, afa, afb in triangle_edges(sk, screen_pct, min_af_pct, threads)
    ]
    return HEADER + "".join(rows)


def rect_tsv(ref_paths, query_paths, screen_pct=80.0, min_af_pct=15.0, threads=1):
    """dist / search: Ref = reference-list genome, Query = query-list genome; rows grouped by query,
    ANI descending."""
    paths = list(dict.fromkeys(list(ref_paths) + list(query_paths)))
    idx = {p: i for i, p in enumerate(paths)}
    sk = sketch_files(paths, threads)
    names = [s.first_name for s in sk]
    rows = []
    for q in dict.fromkeys(query_paths):
        for r in dict.fromkeys(ref_paths):
            a, b = idx[r], idx[q]
            if a == b:
                continue
            if screen_pct > 0 and not O.screen(sk[a], sk[b], screen_pct / 100.0)[1]:
                continue
            res = O.pair(sk[a], sk[b])
            if res.ani < 0 or max(res.af_a, res.af_b) * 100.0 < min_af_pct:
                continue
            rows.append((b, -res.ani * 100.0, a, res))
    rows.sort(key=lambda t: t[:3])
    out = [
        "%s\t%s\t%.2f\t%.2f\t%.2f\t%s\t%s\n"
        % (paths[a], paths[b], r.ani * 100.0, r.af_a * 100.0, r.af_b * 100.0, names[a], names[b])
        for b, _, a, r in rows
    ]
    return HEADER + "".join(out)
