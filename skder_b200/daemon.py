"""Resident sketch database for `skani search`.

skDER's low_mem_greedy mode launches one `skani search <genome> -d <db>` process per representative
(reference src/skDER/skder.py:116-120) -- thousands of launches, each of which would otherwise create a
CUDA context, read the database from disk, upload it and build its index.  The first `search` against a
database starts this server (one per database directory, per GPU); it keeps the indexed database in HBM
and answers each later search with one sketch + one rectangle (Engine.search), then exits after
SKB_DAEMON_IDLE seconds without requests.  The shim talks to it over a Unix socket.
"""
import hashlib
import json
import os
import sys
import time
from multiprocessing.connection import Client, Listener

IDLE_SECONDS = float(os.environ.get("SKB_DAEMON_IDLE", "900"))


def socket_path(db_dir, device):
    key = hashlib.md5(("%s|%d" % (os.path.realpath(db_dir), device)).encode()).hexdigest()[:20]
    return os.path.join(os.environ.get("SKB_DAEMON_DIR", "/tmp"), "skb_%s.sock" % key)


def request(db_dir, device, msg, timeout=None):
    """Send one request; returns the reply dict or None if no server is listening."""
    path = socket_path(db_dir, device)
    if not os.path.exists(path):
        return None
    try:
        with Client(path, family="AF_UNIX") as conn:
            conn.send(msg)
            if timeout is not None and not conn.poll(timeout):
                return None
            return conn.recv()
    except (ConnectionRefusedError, FileNotFoundError, EOFError, OSError):
        return None


def serve(db_dir, device):
    from . import cli, engine

    path = socket_path(db_dir, device)
    if os.path.exists(path):
        if request(db_dir, device, {"op": "ping"}, timeout=2.0):
            return 0  # somebody else already serves this database
        os.unlink(path)
    with open(os.path.join(db_dir, "manifest.json")) as f:
        man = json.load(f)
    paths, names = list(man["paths"]), list(man["names"])
    eng = engine.Engine(device)
    eng.load(db_dir)
    eng.index()
    n_db = eng.n_genomes
    listener = Listener(path, family="AF_UNIX")
    listener._listener._socket.settimeout(1.0)
    last = time.time()
    try:
        while time.time() - last < IDLE_SECONDS:
            try:
                conn = listener.accept()
            except (TimeoutError, OSError):
                continue
            with conn:
                try:
                    msg = conn.recv()
                    if msg.get("op") == "ping":
                        conn.send({"ok": True, "n": n_db})
                    elif msg.get("op") == "stop":
                        conn.send({"ok": True})
                        break
                    elif msg.get("op") == "search":
                        q = msg["query"]
                        packed = engine.pack_fasta(q, eng.params.min_contig_len)
                        edges, st = eng.search(packed, screen=msg["screen"], min_af=msg["min_af"])
                        rows = cli.rect_rows(paths + [q], names + [packed.first_name], edges)
                        cli.write_atomic(msg["out"], cli.HEADER + "".join(rows))
                        conn.send({"ok": True, "rows": len(rows), "ms": st.ms_total})
                    else:
                        conn.send({"ok": False, "error": "unknown op"})
                except Exception as e:  # the shim turns this into "no output file"
                    try:
                        conn.send({"ok": False, "error": "%s: %s" % (type(e).__name__, e)})
                    except Exception:
                        pass
            last = time.time()
    finally:
        listener.close()
        if os.path.exists(path):
            os.unlink(path)
        eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(serve(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0))
