"""Binary edge hand-off to skDER's greedy selection.

reference src/skDER/skder.py:136-148 (greedyDerep) shells out to `skDERsum <edge TSV> <N50 TSV> <ANI> <AF>`, which
parses both text files into std::map<string, ...> keyed by file path (src/skDER/skDERsum.cpp:86-165), and then to
`sort`.  With the edges still binary (skb_edge[] on the device or the host) the same two files are produced without the
text round trip: connectivity and member lists come from skb_greedy_summary (device: exact 2-decimal rounding, counts,
one key sort), this module multiplies by N50, formats and sorts.  Output is byte-identical to the reference helper's.
"""
import os
import subprocess


def _fmt_score(x):
    """C++ `ostream << double` (skDERsum.cpp doubleToString): %g with 6 significant digits."""
    return "%g" % x


def greedy_information(eng, edges, paths_sorted, n50_rows, min_ani, min_af):
    """Text of Genome_Information_for_Greedy_Clustering.txt.
    edges: EDGE_DTYPE array with ids indexing paths_sorted (None: the engine's last device-resident result);
    n50_rows: [(path, n50)] in the order of Concatenated_N50.txt (the order skDERsum prints in)."""
    conn, off, mem = eng.greedy_summary(edges, len(paths_sorted), min_ani, min_af)
    idx = {p: i for i, p in enumerate(paths_sorted)}
    out = []
    for path, n50 in n50_rows:
        g = idx.get(path)
        if g is None or conn[g] == 0:
            out.append(path + "\t0.0\t\n")
        else:
            members = "; ".join(paths_sorted[m] for m in mem[off[g]:off[g + 1]].tolist())
            out.append(path + "\t" + _fmt_score(float(int(n50)) * float(conn[g])) + "\t" + members + "\n")
    return "".join(out)


def write_greedy_information(eng, edges, paths_sorted, n50_file, min_ani, min_af, outdir, threads=1):
    """Both files greedyDerep's loop reads (reference skder.py:141-148): the summary and its `sort -k 2 -gr` order."""
    rows = []
    with open(n50_file) as f:
        for line in f:
            if line.strip():
                p, v = line.rstrip("\n").split("\t")[:2]
                rows.append((p, int(v)))
    summary = os.path.join(outdir, "Genome_Information_for_Greedy_Clustering.txt")
    with open(summary, "w") as f:
        f.write(greedy_information(eng, edges, paths_sorted, rows, min_ani, min_af))
    srt = os.path.join(outdir, "Genome_Information_for_Greedy_Clustering.sorted.txt")
    with open(srt, "w") as f:  # the reference's own sort command (skder.py:146), same locale handling
        subprocess.check_call(["sort", "-k", "2", "--parallel=%d" % threads, "-gr", summary], stdout=f, env=dict(os.environ, LC_ALL="C"))
    return summary, srt
