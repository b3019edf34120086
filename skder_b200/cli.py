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
    out = []
    for e in edges:
        a, b = int(e["a"]), int(e["b"])
        out.append(fmt_row(paths_sorted[a], paths_sorted[b], e["ani"], e["af_a"], e["af_b"], names[a], names[b]))
    return out


def rect_rows(paths, names, edges):
    """edges: a = reference id, b = query id.  Rows grouped by query, ANI descending."""
    rows = sorted(((int(e["b"]), -float(e["ani"]), int(e["a"]), e) for e in edges), key=lambda t: t[:3])
    return [fmt_row(paths[a], paths[b], e["ani"], e["af_a"], e["af_b"], names[a], names[b]) for b, _, a, e in rows]


def _device():
    return int(os.environ.get("SKB_DEVICE", os.environ.get("LOCAL_RANK", "0")))


def run_triangle(opt, engine_mod):
    paths = sorted(set(read_list(opt["list"])))
    with engine_mod.Engine(_device()) as eng:
        packed = eng.add_fasta(paths, threads=opt["threads"])
        names = [p.first_name for p in packed]
        del packed
        eng.index()
        edges, st = eng.triangle(screen=opt["screen"], min_af=opt["min_af"])
    write_atomic(opt["out"], HEADER + "".join(triangle_rows(paths, names, edges)))
    return st


def run_sketch(opt, engine_mod):
    paths = read_list(opt["list"])
    os.makedirs(opt["out"], exist_ok=True)
    with engine_mod.Engine(_device()) as eng:
        packed = eng.add_fasta(paths, threads=opt["threads"])
        names = [p.first_name for p in packed]
        eng.save(opt["out"])
    write_atomic(os.path.join(opt["out"], "manifest.json"), json.dumps({"paths": paths, "names": names}))


def run_search(opt, engine_mod):
    """`skani search q.fa -d DB -o OUT` (reference src/skDER/skder.py:119).  Served by the database's resident
    daemon (skder_b200/daemon.py; started here on first use) unless SKB_NO_DAEMON=1, in which case the database is
    loaded, indexed and searched in this process."""
    query = opt["positional"][0]
    dev = _device()
    if os.environ.get("SKB_NO_DAEMON") != "1":
        import subprocess

        from . import daemon

        msg = {"op": "search", "query": query, "screen": opt["screen"], "min_af": opt["min_af"], "out": opt["out"]}
        rep = daemon.request(opt["db"], dev, msg)
        if rep is None:
            log = open(os.path.join(opt["db"], "skani_b200_daemon.log"), "a")
            subprocess.Popen([sys.executable, "-m", "skder_b200.daemon", opt["db"], str(dev)], stdout=log, stderr=log,
                             stdin=subprocess.DEVNULL, start_new_session=True,
                             cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            deadline = time.time() + float(os.environ.get("SKB_DAEMON_START_TIMEOUT", "600"))
            while rep is None and time.time() < deadline:
                time.sleep(0.2)
                if daemon.request(opt["db"], dev, {"op": "ping"}, timeout=5.0):
                    rep = daemon.request(opt["db"], dev, msg)
        if rep is None:
            raise RuntimeError("search daemon did not come up (see %s/skani_b200_daemon.log)" % opt["db"])
        if not rep.get("ok"):
            raise RuntimeError("search daemon: %s" % rep.get("error"))
        return None
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
        packed = eng.add_fasta(paths, threads=opt["threads"])
        names = [p.first_name for p in packed]
        del packed
        eng.index()
        edges, st = eng.rect([idx[p] for p in dict.fromkeys(refs)], [idx[p] for p in dict.fromkeys(queries)],
                             screen=opt["screen"], min_af=opt["min_af"])
    write_atomic(opt["out"], HEADER + "".join(rect_rows(paths, names, edges)))
    return st


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    out_hint = None
    t0 = time.time()
    try:
        if "-o" in argv[:-1]:
            out_hint = argv[argv.index("-o") + 1]
        sub, opt = parse_args(argv)
        from . import engine as engine_mod  # loads libskani_b200.so; raises if missing (no CPU path)

        {"triangle": run_triangle, "sketch": run_sketch, "search": run_search, "dist": run_dist}[sub](opt, engine_mod)
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
