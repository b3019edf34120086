"""`skani`-compatible command line in front of the B200 engine.

skDER reaches its all-vs-all step only through `subprocess.call('skani ...', shell=True)`
(reference src/skDER/util.py:636-652); putting skder_b200/bin/skani first on PATH swaps the engine
without touching skDER.  Sub-commands and flags are exactly the ones skDER spells:

  skani triangle -l LIST --min-af A -E [-s S] -t T -o OUT        (src/skDER/skder.py:16-18,23-25)
  skani sketch   -l LIST -o DBDIR -t T                           (skder.py:103)
  skani search   QUERY.fa -d DBDIR -o OUT -t T                   (skder.py:119)
  skani dist     --rl REFS --ql QUERIES [-s S] [-t T] -o OUT     (skder.py:58-59, cidder.py:362-363)

Any other skani option (-c, -m, --fast, --slow, --medium, --small-genomes, --no-learned-ani, --robust,
--median, ...) changes skani's estimator in ways this engine does not implement: the shim then exits
non-zero WITHOUT creating the output, which is the one failure signal runCmd checks (util.py:642-645).
stderr is discarded by skDER, so diagnostics also go to <output>.skani_b200.log.
"""
import json
import os
import sys
import time

HEADER = "Ref_file\tQuery_file\tANI\tAlign_fraction_ref\tAlign_fraction_query\tRef_name\tQuery_name\n"
DEFAULT_SCREEN = 80.0  # skani -s default
DEFAULT_MIN_AF = 15.0  # skani --min-af default


class UsageError(Exception):
    pass


def parse_args(argv):
    """Minimal, strict parser for the skani spellings above.  Returns (subcommand, options dict)."""
    if not argv:
        raise UsageError("no sub-command")
    sub = argv[0]
    if sub not in ("triangle", "sketch", "search", "dist"):
        raise UsageError("unsupported sub-command %r" % sub)
    opt = {"screen": DEFAULT_SCREEN, "min_af": DEFAULT_MIN_AF, "threads": 3, "edge_list": False, "positional": []}
    valued = {
        "-l": "list", "-o": "out", "-t": "threads", "-s": "screen", "--min-af": "min_af", "--rl": "rl", "--ql": "ql",
        "-d": "db",
    }
    flags = {"-E": "edge_list", "--sparse": "edge_list"}
    i = 1
    while i < len(argv):
        a = argv[i]
        if a == "":
            i += 1
            continue
        if a in valued:
            if i + 1 >= len(argv):
                raise UsageError("option %s needs a value" % a)
            opt[valued[a]] = argv[i + 1]
            i += 2
        elif a in flags:
            opt[flags[a]] = True
            i += 1
        elif a.startswith("-"):
            raise UsageError("skani option %r is not implemented by the B200 engine" % a)
        else:
            opt["positional"].append(a)
            i += 1
    try:
        opt["screen"] = float(opt["screen"])
        opt["min_af"] = float(opt["min_af"])
        opt["threads"] = max(1, int(opt["threads"]))
    except ValueError as e:
        raise UsageError("bad numeric option: %s" % e)
    need = {"triangle": ["list", "out"], "sketch": ["list", "out"], "search": ["db", "out"], "dist": ["rl", "ql", "out"]}[sub]
    for k in need:
        if k not in opt:
            raise UsageError("skani %s: missing required option for %r" % (sub, k))
    if sub == "triangle" and not opt["edge_list"]:
        raise UsageError("skani triangle without -E (matrix output) is not implemented; skDER always passes -E")
    if sub == "search" and len(opt["positional"]) < 1:
        raise UsageError("skani search: no query FASTA given")
    if sub != "search" and opt["positional"]:
        raise UsageError("unexpected positional arguments: %r" % opt["positional"])
    return sub, opt


def read_list(path):
    with open(path) as f:
        return [ln.strip() for ln in f if ln.strip()]


def fmt_row(ref, query, ani, af_ref, af_query, ref_name, query_name):
    return "%s\t%s\t%.2f\t%.2f\t%.2f\t%s\t%s\n" % (ref, query, ani, af_ref, af_query, ref_name, query_name)


def write_atomic(path, text):
    d = os.path.dirname(os.path.abspath(path))
    tmp = os.path.join(d, ".%s.tmp.%d" % (os.path.basename(path), os.getpid()))
    with open(tmp, "w") as f:
        f.write(text)
    os.replace(tmp, path)


def triangle_rows(paths_sorted, names, edges):
    """edges: structured array (a < b, ids index paths_sorted).  Ref = lexicographically smaller path
    (SURVEY.md section 4 fact 3); rows grouped by Ref."""
    cols = [edges[k].tolist() for k in ("a", "b", "ani", "af_a", "af_b")]  # element access on structured arrays is slow
    return [fmt_row(paths_sorted[a], paths_sorted[b], ani, afa, afb, names[a], names[b]) for a, b, ani, afa, afb in zip(*cols)]


def rect_rows(paths, names, edges):
    """edges: a = reference id, b = query id.  Rows grouped by query, ANI descending."""
    cols = [edges[k].tolist() for k in ("a", "b", "ani", "af_a", "af_b")]
    rows = sorted(zip(*cols), key=lambda t: (t[1], -t[2], t[0]))
    return [fmt_row(paths[a], paths[b], ani, afa, afb, names[a], names[b]) for a, b, ani, afa, afb in rows]


def _device():
    return int(os.environ.get("SKB_DEVICE", os.environ.get("LOCAL_RANK", "0")))


INGEST_BATCH = 256  # genomes packed per host batch: batch k+1 is parsed while batch k uploads and is sketched


def stream_add(eng, engine_mod, paths, threads, phases=None):
    """FASTA files -> sketches on the device, streamed: host threads parse and 2-bit pack batch k+1
    (skb_pack_fasta_many) while the device uploads and sketches batch k (skb_add_genomes), so at most two batches of
    packed genomes are resident on the host and the ingest hides behind the upload.  Returns the first-record names."""
    from concurrent.futures import ThreadPoolExecutor

    names, t_wait, t_add = [], 0.0, 0.0
    mcl = eng.params.min_contig_len
    with ThreadPoolExecutor(1) as ex:
        fut = ex.submit(engine_mod.pack_fasta_many, paths[:INGEST_BATCH], threads, mcl) if paths else None
        for i in range(0, len(paths), INGEST_BATCH):
            t0 = time.time()
            packed = fut.result()
            nxt = paths[i + INGEST_BATCH:i + 2 * INGEST_BATCH]
            fut = ex.submit(engine_mod.pack_fasta_many, nxt, threads, mcl) if nxt else None
            t1 = time.time()
            names += [p.first_name for p in packed]
            eng.add(packed)
            del packed
            t_wait += t1 - t0
            t_add += time.time() - t1
    if phases is not None:
        phases["ingest_wait_s"] = t_wait  # time the device side waited for the parser (first batch + any shortfall)
        phases["upload_sketch_s"] = t_add
    return names


def _phase_log(opt, phases):
    if os.environ.get("SKB_PHASE_LOG") == "1":
        write_atomic(opt["out"].rstrip("/") + ".skani_b200.phases.json", json.dumps(phases))


def run_triangle(opt, engine_mod):
    ph, t0 = {}, time.time()
    paths = sorted(set(read_list(opt["list"])))
    with engine_mod.Engine(_device()) as eng:
        ph["context_s"] = time.time() - t0
        names = stream_add(eng, engine_mod, paths, opt["threads"], ph)
        t1 = time.time()
        eng.index()
        t2 = time.time()
        edges, st = eng.triangle(screen=opt["screen"], min_af=opt["min_af"])
        t3 = time.time()
    write_atomic(opt["out"], HEADER + "".join(triangle_rows(paths, names, edges)))
    ph.update({"index_s": t2 - t1, "triangle_s": t3 - t2, "write_tsv_s": time.time() - t3, "total_s": time.time() - t0,
               "genomes": len(paths), "edges": int(len(edges))})
    _phase_log(opt, ph)
    return st


def run_sketch(opt, engine_mod):
    paths = read_list(opt["list"])
    os.makedirs(opt["out"], exist_ok=True)
    from . import daemon

    daemon.stop_for(opt["out"])  # a server still holding the previous contents of this directory must not answer for the new ones
    with engine_mod.Engine(_device()) as eng:
        names = stream_add(eng, engine_mod, paths, opt["threads"])
        eng.save(opt["out"])
    write_atomic(os.path.join(opt["out"], "manifest.json"), json.dumps({"paths": paths, "names": names}))


def run_search(opt, engine_mod):
    """`skani search q.fa -d DB -o OUT` (reference src/skDER/skder.py:119).  Served by the database's resident
    daemon (skder_b200/daemon.py; started here on first use) unless SKB_NO_DAEMON=1, in which case the database is
    loaded, indexed and searched in this process."""
    # the server runs in another working directory: every path goes over absolute
    query, db, out = os.path.abspath(opt["positional"][0]), os.path.abspath(opt["db"]), os.path.abspath(opt["out"])
    opt = dict(opt, db=db, out=out)
    dev = _device()
    if os.environ.get("SKB_NO_DAEMON") != "1":
        import subprocess

        from . import daemon

        msg = {"op": "search", "query": query, "label": opt["positional"][0],  # the row shows the path as it was given
               "screen": opt["screen"], "min_af": opt["min_af"], "out": out}
        rep = daemon.request(db, dev, msg)
        if rep is not None and rep.get("stale"):  # the directory was rewritten under a running server: it has just exited
            deadline = time.time() + 15.0
            while time.time() < deadline and daemon.request(db, dev, {"op": "ping"}, timeout=1.0) is not None:
                time.sleep(0.05)
            rep = None
        if rep is None:
            log = open(os.path.join(db, "skani_b200_daemon.log"), "a")
            subprocess.Popen([sys.executable, "-m", "skder_b200.daemon", db, str(dev)], stdout=log, stderr=log,
                             stdin=subprocess.DEVNULL, start_new_session=True,
                             cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            deadline = time.time() + float(os.environ.get("SKB_DAEMON_START_TIMEOUT", "600"))
            while rep is None and time.time() < deadline:
                time.sleep(0.2)
                if daemon.request(db, dev, {"op": "ping"}, timeout=5.0):
                    rep = daemon.request(db, dev, msg)
        if rep is None:
            raise RuntimeError("search daemon did not come up (see %s/skani_b200_daemon.log)" % db)
        if not rep.get("ok"):
            raise RuntimeError("search daemon: %s" % rep.get("error"))
        return None
    if engine_mod is None:
        from . import engine as engine_mod
    query = opt["positional"][0]
    with open(os.path.join(opt["db"], "manifest.json")) as f:
        man = json.load(f)
    paths, names = list(man["paths"]), list(man["names"])
    with engine_mod.Engine(dev) as eng:
        eng.load(opt["db"])
        eng.index()
        qp = engine_mod.pack_fasta(query, eng.params.min_contig_len)
        edges, st = eng.search(qp, screen=opt["screen"], min_af=opt["min_af"])
    write_atomic(opt["out"], HEADER + "".join(rect_rows(paths + [query], names + [qp.first_name], edges)))
    return st


def run_dist(opt, engine_mod):
    refs, queries = read_list(opt["rl"]), read_list(opt["ql"])
    paths = list(dict.fromkeys(refs + queries))
    idx = {p: i for i, p in enumerate(paths)}
    with engine_mod.Engine(_device()) as eng:
        names = stream_add(eng, engine_mod, paths, opt["threads"])
        eng.index()
        edges, st = eng.rect([idx[p] for p in dict.fromkeys(refs)], [idx[p] for p in dict.fromkeys(queries)],
                             screen=opt["screen"], min_af=opt["min_af"])
    write_atomic(opt["out"], HEADER + "".join(rect_rows(paths, names, edges)))
    return st


PROVENANCE = ("skani %s: produced by skder_b200 (B200 engine) in %.2f s -- NOT the skani binary: a restatement of the published "
              "skani method (Shaw & Yu 2023) whose ANI/AF agree with skani's on the reference's golden pairs to sd 0.15 pp "
              "ANI / 0.44 pp AF, out of fold (tests/golden/ORACLE_VS_GOLDEN.md); sketches and prescreen decisions are not "
              "pinned against skani\n")


def _provenance(opt, sub, seconds):
    """Say beside the output which estimator wrote it (skDER discards stdout/stderr).  Overwritten, not appended: the
    search output is rewritten thousands of times by the low_mem_greedy loop."""
    try:
        write_atomic(opt["out"].rstrip("/") + ".skani_b200.log", PROVENANCE % (sub, seconds))
    except OSError:
        pass


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    out_hint = None
    t0 = time.time()
    try:
        if "-o" in argv[:-1]:
            out_hint = argv[argv.index("-o") + 1]
        sub, opt = parse_args(argv)
        # A `search` the database's daemon answers needs neither numpy nor the CUDA library in THIS process: skDER's
        # low_mem_greedy loop launches one `skani search` per representative (src/skDER/skder.py:116-120), and importing
        # the engine costs ~0.5 s a time against ~0.1 s for the interpreter plus the socket client.
        engine_mod = None
        if not (sub == "search" and os.environ.get("SKB_NO_DAEMON") != "1"):
            from . import engine as engine_mod  # loads libskani_b200.so; raises if missing (no CPU path)

        {"triangle": run_triangle, "sketch": run_sketch, "search": run_search, "dist": run_dist}[sub](opt, engine_mod)
        _provenance(opt, sub, time.time() - t0)
        return 0
    except Exception as e:  # no output file is left behind: skDER's runCmd then raises
        msg = "skani (B200 engine) failed after %.1fs: %s: %s\nargv: %r\n" % (time.time() - t0, type(e).__name__, e, argv)
        sys.stderr.write(msg)
        if out_hint:
            try:
                with open(out_hint.rstrip("/") + ".skani_b200.log", "a") as f:
                    f.write(msg)
            except OSError:
                pass
        return 2


if __name__ == "__main__":
    sys.exit(main())
